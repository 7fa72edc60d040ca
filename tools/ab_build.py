"""Tuning aid: build a variant of libqocgrape.so with extra nvcc flags (e.g. -DQOC_EXPM_MINB=6) into build/variants/<name>.so;
run it with QOCGRAPE_LIB=<path>.   python tools/ab_build.py <name> <unit.cu> [flags...]   (only <unit.cu> is recompiled)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quoptimalcontrol.jl_b200"))
import build as b
name, unit, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
b.build()
vdir = os.path.join(b.HERE, "build", "variants")
os.makedirs(vdir, exist_ok=True)
obj = os.path.join(vdir, name + ".o")
subprocess.run(["nvcc"] + b.NVCC_FLAGS + flags + ["-c", "-o", obj, os.path.join(b.CSRC, unit)], check=True)
objs = [obj if s == unit else b._obj(s) for s in b.SOURCES]
out = os.path.join(vdir, name + ".so")
subprocess.run(["nvcc", "-shared", "-o", out] + objs, check=True)
print(out)
