"""Builds libqocgrape.so (CUDA kernels + C ABI) in-tree for sm_100a with nvcc.  No torch involved.

The library is four translation units compiled in parallel into build/*.o and linked with nvcc: the host
orchestration + C ABI (qocgrape.cu) and three kernel families (fused warp-per-chain, phased / chunk-parallel,
large-dimension GEMM pipeline).  An object is rebuilt when its source or one of the headers it includes is newer."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libqocgrape.so")
ABI = os.path.join("..", "..", "include", "qocgrape.h")
UNITS = {   # source -> headers it depends on
    "qocgrape.cu": ["params.h", "big_api.h", ABI],
    "k_small_fused.cu": ["params.h", "warp_mat.cuh", "small_d.cuh", "common_kernels.cuh"],
    "k_small_phased.cu": ["params.h", "warp_mat.cuh", "small_d.cuh", "small_phased.cuh"],
    "k_big.cu": ["params.h", "big_api.h", "big_d.cuh", "zgemm_dmma.cuh", "pure_state.cuh", ABI],
}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
SOURCES = list(UNITS)


def _obj(src):
    return os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")


def _unit_stale(src):
    o = _obj(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    deps = [os.path.join(CSRC, d) for d in [src] + UNITS[src]] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _stale():
    if not os.path.exists(LIB):
        return True
    return any(_unit_stale(s) or os.path.getmtime(_obj(s)) > os.path.getmtime(LIB) for s in SOURCES)


def build(force=False, verbose=False, extra_flags=()):
    """Compile what is out of date and link.  Returns the library path."""
    if not force and not extra_flags and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libqocgrape.so cannot be built (and there is no CPU fallback)")
    os.makedirs(OBJDIR, exist_ok=True)
    todo = [s for s in SOURCES if force or extra_flags or _unit_stale(s)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + \
            ["-c", "-o", _obj(src), os.path.join(CSRC, src)]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:
        for src, res in pool.map(compile_one, todo):
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n" + res.stdout + res.stderr)
            if verbose:
                print(res.stderr)
    res = subprocess.run([nvcc, "-shared", "-o", LIB] + [_obj(s) for s in SOURCES], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
