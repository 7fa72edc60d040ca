"""The C-ABI library loads and exports every symbol include/qocgrape.h declares; struct layouts of the ctypes
binding match the header; without a GPU the product path fails loudly (no CPU fallback)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import quoptimalcontrol_jl_b200 as qoc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qocgrape.h")


def _declared_functions():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(qoc_[a-z_]+)\s*\(", txt)))


def test_every_declared_symbol_is_exported():
    lib = qoc._lib.load()
    names = _declared_functions()
    assert set(names) == set(qoc._lib.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.qoc_version().decode().startswith("qocgrape")


def test_struct_layouts_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "qocgrape.h"\nint main(){printf("%zu %zu %zu %zu\\n",'
                   'sizeof(qoc_desc),sizeof(qoc_stats),offsetof(qoc_desc,T),offsetof(qoc_stats,workspace_bytes));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert int(out[0]) == ctypes.sizeof(qoc._lib.QocDesc)
    assert int(out[1]) == ctypes.sizeof(qoc._lib.QocStats)
    assert int(out[2]) == qoc._lib.QocDesc.T.offset
    assert int(out[3]) == qoc._lib.QocStats.workspace_bytes.offset


def test_header_compiles_as_c_and_cpp(tmp_path):
    for comp, ext in (("gcc", "c"), ("g++", "cpp")):
        f = tmp_path / f"t.{ext}"
        f.write_text('#include "qocgrape.h"\nint main(void){return QOC_OK;}\n')
        subprocess.run([comp, "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(f), "-o", str(tmp_path / "t.o")], check=True)


def test_invalid_descriptor_is_rejected():
    lib = qoc._lib.load()
    h = ctypes.c_void_p()
    d = qoc._lib.QocDesc(sys_type=7, D=2, K=1, N=1, M=1, R=1, T=1.0)
    assert lib.qoc_create(ctypes.byref(h), ctypes.byref(d)) == qoc._lib.QOC_EINVAL
    assert b"invalid" in lib.qoc_last_error(None)


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    Z = np.zeros((2, 2), dtype=complex)
    with pytest.raises(qoc.QocError) as e:
        qoc.GrapeEvaluator([(Z, [Z], Z, Z)], 1.0, 4, qoc._lib.STATE_TRANSFER)
    assert e.value.status == qoc._lib.QOC_ECUDA and "no CPU fallback" in str(e.value)


def test_product_path_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "quoptimalcontrol.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt, os.path.join(dirpath, f)
