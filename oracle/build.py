"""Builds the C restatement (oracle/grape_oracle.c) into oracle/libqoc_oracle{,_avx2}.so with gcc.
TEST INFRASTRUCTURE: the checker / reported CPU baseline, never part of the product path."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "grape_oracle.c")
VARIANTS = {"libqoc_oracle.so": [], "libqoc_oracle_avx2.so": ["-mavx2", "-mfma"]}


def build(force=False):
    out = []
    for name, extra in VARIANTS.items():
        lib = os.path.join(HERE, name)
        if force or not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(SRC):
            cmd = ["gcc", "-O3", "-std=c11", "-fPIC", "-shared", "-fopenmp", "-fno-math-errno", "-fcx-limited-range"] + extra + \
                ["-o", lib, SRC, "-lm"]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError("gcc failed:\n" + res.stderr)
        out.append(lib)
    return out


if __name__ == "__main__":
    print(build(force=True))
