"""Generates tests/golden/cfg5_full.npz: F and G[16, 2000] of the FULL-SIZE BASELINE config 5 (8-qubit Ising chain,
D = 256, K = 16, N = 2000, literal |0...0> -> |1...1> instance, seed 1005) from the CPU oracle.

The oracle call is the literal reference loop order (oracle/grape_oracle.py::fom_and_gradient_grape: expm per slice,
two GEMMs per slice and direction, three GEMMs per (control, slice): ~100 000 complex 256^3 GEMMs), run ONCE offline
(about 5 minutes on 8 host cores with OpenBLAS); the GPU tests and bench.py then check the full-size evaluation
-- including its Kogge-Stone chunk-boundary products, which no reduced-size test reaches -- against this file.
A second, independent evaluation of the same gradient through the trace identity
    g[c,t] = Re tr( i dt C_t' [B_c, S_t] ) = Re( i dt * sum(B_c^T .* (S_t C_t' - C_t' S_t)) )
is computed alongside and must agree to 1e-12 (guards against a slip in either form).

`--variant dense` writes cfg5_full_dense.npz: same operators, pulse and sizes, but seeded dense random density
matrices as Xi / Xt (`dense_states`, seed 99).  The literal instance barely moves |0...0> towards |1...1> (F = 1 to all
16 digits, |G| ~ 2e-13: the parity check on it is purely relative); the dense variant has F and G of order 1e-2.

Like every fixture here it comes from the oracle, not from a Julia run (Julia is not installed): PARITY UNPINNED.
Run from the repo root:  python tests/golden/make_golden_cfg5.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import grape_oracle as orc  # noqa: E402


def load_configs():
    """configs.py is numpy-only; load it without importing the package (which would want the CUDA library)."""
    import importlib.util
    import types
    pkg = types.ModuleType("_qoc_cfg_pkg")
    pkg.__path__ = [os.path.join(ROOT, "quoptimalcontrol.jl_b200")]
    sys.modules["_qoc_cfg_pkg"] = pkg
    lib = types.ModuleType("_qoc_cfg_pkg._lib")
    lib.STATE_TRANSFER, lib.UNITARY_GATE, lib.COHERENCE_TRANSFER = 0, 1, 2
    sys.modules["_qoc_cfg_pkg._lib"] = lib
    spec = importlib.util.spec_from_file_location("_qoc_cfg_pkg.configs", os.path.join(ROOT, "quoptimalcontrol.jl_b200", "configs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def dense_states(D, seed=99):
    """Seeded dense random density matrices (positive, trace 1) used as Xi / Xt by the `dense` variant."""
    rng = np.random.default_rng(seed)

    def dens():
        Z = rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))
        Z = Z @ Z.conj().T
        return Z / np.trace(Z).real
    return dens(), dens()


def main():
    variant = "dense" if "dense" in sys.argv[1:] or "--variant=dense" in sys.argv[1:] else "literal"
    cfg = load_configs().config5()
    A, B, Xi, Xt = cfg["members"][0]
    if variant == "dense":
        Xi, Xt = dense_states(A.shape[0])
    x, T, N = cfg["x"], cfg["T"], cfg["N"]
    K = x.shape[0]
    dt = T / N
    t0 = time.time()
    F, G, P, S, C = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, cfg["sys_type"], return_stores=True)
    t1 = time.time()
    print(f"reference-order evaluation: {t1 - t0:.1f} s   F = {F!r}   |G|_inf = {np.max(np.abs(G)):.3e}", flush=True)
    G2 = np.zeros_like(G)
    BT = [b.T.copy() for b in B]
    for t in range(N):
        W = S[t] @ orc.dag(C[t]) - orc.dag(C[t]) @ S[t]
        for c in range(K):
            G2[c, t] = np.real(1.0j * dt * np.sum(BT[c] * W))
    dev = np.max(np.abs(G2 - G)) / np.max(np.abs(G))
    print(f"trace-identity evaluation: {time.time() - t1:.1f} s   max deviation / |G|_inf = {dev:.3e}", flush=True)
    assert dev < 1e-12
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg5_full.npz" if variant == "literal" else "cfg5_full_dense.npz")
    np.savez_compressed(path, F=np.array(F), G=G, x=x, meta=np.array([T, N, K, A.shape[0], 1005], dtype=np.float64))
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
