"""Generates tests/golden/grape_golden.npz: inputs and oracle outputs of small GRAPE cases.

The reference (Julia) cannot be executed in this environment and ships no fixtures, so these vectors come from the
CPU oracle (oracle/grape_oracle.py, itself pinned by analytic known answers, mpmath and finite differences in
tests/test_oracle.py).  They freeze the oracle against drift and give the GPU tests a second, file-based checker.
Run from the repo root:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import grape_oracle as orc  # noqa: E402
from conftest import random_system  # noqa: E402

CASES = [  # name, D, K, N, M, T, sys_type, mode
    ("state_d2", 2, 2, 10, 1, 1.0, orc.STATE_TRANSFER, "inplace"),
    ("unitary_d2_static", 2, 2, 10, 1, 1.0, orc.UNITARY_GATE, "static"),
    ("unitary_d4_exact", 4, 4, 12, 1, 2.0, orc.UNITARY_GATE, "exact"),
    ("coherence_d16", 16, 3, 6, 1, 0.8, orc.COHERENCE_TRANSFER, "inplace"),
    ("ensemble_d8", 8, 6, 9, 5, 1.5, orc.UNITARY_GATE, "inplace"),
    ("state_d8_exact_ens", 8, 2, 7, 3, 1.0, orc.STATE_TRANSFER, "exact"),
    ("state_d64", 64, 2, 5, 1, 0.5, orc.STATE_TRANSFER, "inplace"),
]


def build(name, D, K, N, M, T, sys_type, mode, seed):
    members = [random_system(D, K, seed=seed + k, hermitian=(sys_type != orc.COHERENCE_TRANSFER),
                             unitary_targets=(sys_type == orc.UNITARY_GATE)) for k in range(M)]
    wts = np.linspace(0.5, 1.0, M) / M
    x = np.random.default_rng(seed).uniform(-1, 1, (K, N))
    if mode == "exact":
        F, G = orc.ensemble_exact(members, wts, x, T, sys_type)
    else:
        F, G = orc.ensemble_fom_and_gradient(members, wts, x, T, sys_type,
                                             orc.REF_STATIC if mode == "static" else orc.REF_INPLACE)
    return members, wts, x, F, G


def main():
    out = {}
    for i, (name, D, K, N, M, T, st, mode) in enumerate(CASES):
        members, wts, x, F, G = build(name, D, K, N, M, T, st, mode, 5000 + 10 * i)
        out[name + "/A"] = np.array([m[0] for m in members])
        out[name + "/B"] = np.array([m[1] for m in members])
        out[name + "/Xi"] = np.array([m[2] for m in members])
        out[name + "/Xt"] = np.array([m[3] for m in members])
        out[name + "/wts"] = wts
        out[name + "/x"] = x
        out[name + "/meta"] = np.array([T, st, {"inplace": 0, "static": 1, "exact": 2}[mode]], dtype=np.float64)
        out[name + "/F"] = np.array(F)
        out[name + "/G"] = G
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "grape_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
