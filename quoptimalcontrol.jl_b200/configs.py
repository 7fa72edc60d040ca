"""The five BASELINE.json workloads as synthetic Hamiltonians / pulses, plus scaled-down
variants for parity tests.  numpy only; seeds fixed so every consumer (tests, bench, CPU checker) sees identical bits."""
from __future__ import annotations

from functools import reduce

import numpy as np

from . import _lib

I2 = np.eye(2, dtype=np.complex128)
SX = np.array([[0, 1], [1, 0]], dtype=np.complex128)
SY = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
SZ = np.array([[1, 0], [0, -1]], dtype=np.complex128)


def kron_all(ops):
    return reduce(np.kron, ops)


def op_on(op, q, n):
    """`op` on qubit q (0-based) of n qubits."""
    return kron_all([op if i == q else I2 for i in range(n)])


def h_super(H):
    """Column-stacking Liouville superoperator of the commutator: Hs(H) = I (x) H - H^T (x) I."""
    D = H.shape[0]
    return np.kron(np.eye(D), H) - np.kron(H.T, np.eye(D))


def config1(variant="test", seed=1001, N=10):
    """Single-qubit StateTransfer |0> -> |1>, controls sx, sy, 10 slices, T = 1
    (/root/reference/test/state_transfer_tests.jl:6-17; README.md:12-24 for variant="readme")."""
    rng = np.random.default_rng(seed)
    if variant == "readme":
        B, A = [SX, SY], 0.0 * SZ
    else:
        B, A = [SX / 2, SY / 2], SZ / 2
    rho0 = np.array([[1, 0], [0, 0]], dtype=np.complex128)
    rho1 = np.array([[0, 0], [0, 1]], dtype=np.complex128)
    return dict(name="cfg1-single-qubit-state-transfer", members=[(A, B, rho0, rho1)], wts=None, T=1.0, N=N,
                sys_type=_lib.STATE_TRANSFER, gradient="first_order", x=rng.random((2, N)))


def _two_qubit():
    A = 0.5 * np.kron(SZ, SZ)
    B = [np.kron(SX / 2, I2), np.kron(SY / 2, I2), np.kron(I2, SX / 2), np.kron(I2, SY / 2)]
    return A, B


def config2(seed=1002, N=1000, T=40.0):
    """Two-qubit UnitarySynthesis CNOT, 4 controls, 1000 slices, exact gradient."""
    rng = np.random.default_rng(seed)
    A, B = _two_qubit()
    cnot = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=np.complex128)
    return dict(name="cfg2-two-qubit-cnot-exact", members=[(A, B, np.eye(4, dtype=np.complex128), cnot)], wts=None,
                T=T * N / 1000.0, N=N, sys_type=_lib.UNITARY_GATE, gradient="exact", x=rng.uniform(-1, 1, (4, N)))


def config3(seed=1003, N=2000, T=40.0, gamma=0.01):
    """OpenSystemCoherenceTransfer: two-qubit Liouvillian (dim 16, pure dephasing), 2000 slices.
    Reference convention P = exp(-i dt A)  =>  A = Hs(H0) + i*gamma*sum_q (conj(Z_q) (x) Z_q - I)."""
    rng = np.random.default_rng(seed)
    H0, Bh = _two_qubit()
    A = h_super(H0).astype(np.complex128)
    for q in range(2):
        Z = op_on(SZ, q, 2)
        A = A + 1j * gamma * (np.kron(Z.conj(), Z) - np.eye(16))
    B = [h_super(b) for b in Bh]
    r0 = np.zeros((4, 4), dtype=np.complex128); r0[0, 0] = 1
    rt = np.zeros((4, 4), dtype=np.complex128); rt[3, 3] = 1
    v0, vt = r0.reshape(-1, order="F"), rt.reshape(-1, order="F")
    Xi, Xt = np.outer(v0, v0.conj()), np.outer(vt, vt.conj())
    return dict(name="cfg3-two-qubit-liouvillian", members=[(A, B, Xi, Xt)], wts=None, T=T * N / 2000.0, N=N,
                sys_type=_lib.COHERENCE_TRANSFER, gradient="first_order", x=rng.uniform(-1, 1, (4, N)))


def config4(seed=1004, N=500, T=10.0, grid=64):
    """Robust GRAPE ensemble: grid x grid detuning/amplitude samples of a 3-qubit gate (Toffoli), 500 slices."""
    rng = np.random.default_rng(seed)
    n = 3
    A0 = 0.5 * (kron_all([SZ, SZ, I2]) + kron_all([I2, SZ, SZ]))
    Sz_tot = sum(op_on(SZ / 2, q, n) for q in range(n))
    B0 = []
    for q in range(n):
        B0 += [op_on(SX / 2, q, n), op_on(SY / 2, q, n)]
    toff = np.eye(8, dtype=np.complex128)
    toff[6:, 6:] = np.array([[0, 1], [1, 0]])
    Xi = np.eye(8, dtype=np.complex128)
    deltas = np.linspace(-0.2, 0.2, grid) if grid > 1 else np.array([0.0])
    epss = np.linspace(-0.05, 0.05, grid) if grid > 1 else np.array([0.0])
    members = []
    for da in deltas:
        for eb in epss:
            members.append((A0 + da * Sz_tot, [(1 + eb) * b for b in B0], Xi, toff))
    M = len(members)
    return dict(name=f"cfg4-robust-ensemble-{M}", members=members, wts=np.full(M, 1.0 / M), T=T * N / 500.0, N=N,
                sys_type=_lib.UNITARY_GATE, gradient="first_order", x=rng.uniform(-1, 1, (6, N)))


def config5(seed=1005, N=2000, T=20.0, n=8):
    """n-qubit Ising spin chain state transfer |0..0> -> |1..1> (dim 2^n), 2n controls, N slices."""
    rng = np.random.default_rng(seed)
    D = 2 ** n
    A = sum(op_on(SZ, q, n) @ op_on(SZ, q + 1, n) for q in range(n - 1))
    B = []
    for q in range(n):
        B += [op_on(SX / 2, q, n), op_on(SY / 2, q, n)]
    Xi = np.zeros((D, D), dtype=np.complex128); Xi[0, 0] = 1
    Xt = np.zeros((D, D), dtype=np.complex128); Xt[-1, -1] = 1
    return dict(name=f"cfg5-ising-{n}q", members=[(A.astype(np.complex128), B, Xi, Xt)], wts=None,
                T=T * N / 2000.0, N=N, sys_type=_lib.STATE_TRANSFER, gradient="first_order",
                x=rng.uniform(-1, 1, (2 * n, N)))


def alg_flops(cfg):
    """Algorithmic FLOPs of one evaluation: M*N*8D^3*(E + P + Gm), E = 3 (the reference's Pade-5
    rung for these inputs), P = 2 / 4 and Gm = 1 / 2 for unitary / density types; exact: E*(1+2K)."""
    A, B = cfg["members"][0][0], cfg["members"][0][1]
    D, K, M, N = A.shape[0], len(B), len(cfg["members"]), cfg["N"]
    unitary = cfg["sys_type"] == _lib.UNITARY_GATE
    E = 3 * (1 + 2 * K) if cfg["gradient"] == "exact" else 3
    return M * N * 8 * D ** 3 * (E + (2 if unitary else 4) + (1 if unitary else 2))
