"""GPU parity of the structure-dependent forms of the D = 5..8 closed-system kernels against the oracle: the plane-wise DMMA
generator assembly (at most 4 matrices with a real plane, 4 with an imaginary plane; natural blocks or compact blocks over the
union of the non-zero entries) and the trace-dots over the union of the controls' non-zero entries.  Both are decided in qoc_set_system from the matrices of ALL members; the cases below flip every
decision (and the environment switches force the dense forms on the same inputs).  Chain counts are chosen so that the
chunk-parallel closed-system kernels run (>= 150 chains, N >= 64)."""
import os

import numpy as np
import pytest

import quoptimalcontrol_jl_b200 as qoc
from oracle import c_oracle
from oracle import grape_oracle as orc
from conftest import assert_parity

pytestmark = pytest.mark.gpu

SX = np.array([[0, 1], [1, 0]], dtype=np.complex128)
SY = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
SZ = np.array([[1, 0], [0, -1]], dtype=np.complex128)
I2 = np.eye(2, dtype=np.complex128)


def _on(op, q, n=3):
    out = np.array([[1.0 + 0j]])
    for i in range(n):
        out = np.kron(out, op if i == q else I2)
    return out


def _members(kind, D, M, seed):
    rng = np.random.default_rng(seed)
    herm = lambda X: (X + X.conj().T) / 2
    if kind == "pauli":            # purely real / purely imaginary controls, diagonal drift (cfg4's structure), D = 8
        A0 = 0.5 * (_on(SZ, 0) @ _on(SZ, 1) + _on(SZ, 1) @ _on(SZ, 2))
        B0 = [_on(SX, 0) / 2, _on(SY, 0) / 2, _on(SX, 1) / 2, _on(SY, 1) / 2, _on(SX, 2) / 2, _on(SY, 2) / 2]
    elif kind == "pauli_k7":       # 7 controls: 4 real + 3 imaginary, complex drift -> real-plane list has 4 entries, imaginary 5 (dense assembly)
        A0 = herm(rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8))) / 3
        B0 = [_on(SX, 0) / 2, _on(SY, 0) / 2, _on(SX, 1) / 2, _on(SY, 1) / 2, _on(SX, 2) / 2, _on(SY, 2) / 2, _on(SZ, 0) / 2]
    elif kind == "real_sparse":    # all controls real and sparse, drift real: the real plane of -i dt X is empty everywhere
        A0 = np.diag(rng.standard_normal(D)).astype(np.complex128)
        B0 = []
        for j in range(3):
            X = np.zeros((D, D), dtype=np.complex128)
            X[j, j + 1] = X[j + 1, j] = 0.7 + 0.1 * j
            B0.append(X)
    elif kind == "mixed_member":   # one member's control has both planes, the others only one: the union decides
        A0 = np.diag(rng.standard_normal(D)).astype(np.complex128)
        X = np.zeros((D, D), dtype=np.complex128); X[0, 1] = X[1, 0] = 0.5
        Y = np.zeros((D, D), dtype=np.complex128); Y[1, 2] = -0.5j; Y[2, 1] = 0.5j
        B0 = [X, Y]
    else:                          # dense complex Hermitian: dense assembly, dense dots
        A0 = herm(rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))) / np.sqrt(D)
        B0 = [herm(rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))) / np.sqrt(D) for _ in range(3)]
    D = A0.shape[0]
    Xi = np.eye(D, dtype=np.complex128)
    Xt = np.linalg.qr(rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D)))[0]
    members = []
    for k in range(M):
        eps, dlt = 0.05 * rng.standard_normal(), 0.2 * rng.standard_normal()
        B = [(1 + eps) * b for b in B0]
        if kind == "mixed_member" and k == M // 2:
            B[0] = B[0] + 0.3 * B0[1]
        members.append((A0 + dlt * np.diag(np.arange(D) - D / 2).astype(np.complex128), B, Xi, Xt))
    return members


@pytest.mark.parametrize("kind,D", [("pauli", 8), ("pauli_k7", 8), ("real_sparse", 8), ("real_sparse", 6), ("mixed_member", 5),
                                    ("dense", 8), ("dense", 7)])
@pytest.mark.parametrize("env", [{}, {"QOC_ASM_COMPACT": "0"}, {"QOC_ASM_SPARSE": "0", "QOC_DOTS_SPARSE": "0"}, {"QOC_PERSIST": "1"}])
def test_structure_dependent_closed_kernels(monkeypatch, kind, D, env):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    M, N, T = 160, 67, 1.7
    members = _members(kind, D, M, seed=11 + D)
    K = len(members[0][1])
    wts = np.random.default_rng(5).uniform(0.5, 1.5, M); wts /= wts.sum()
    x = np.random.default_rng(6).uniform(-1, 1, (K, N))
    Fo, Go = c_oracle.eval_ensemble(members, wts, x, T, orc.UNITARY_GATE, 0, os.cpu_count() or 1)
    with qoc.GrapeEvaluator(members, T, N, orc.UNITARY_GATE, wts=wts) as ev:
        F, G = ev.eval(x)
        assert ev.stats()["launches_last_eval"] >= 3
    assert_parity(F, G, Fo, Go)
