"""CPU tests that pin the oracle (the reference ships no fixtures, see oracle/grape_oracle.py header):
analytic known answers, 50-digit mpmath exponentials, finite differences of the reference's AD functional,
and structural invariants."""
import numpy as np
import pytest

from oracle import grape_oracle as orc
from conftest import random_system

Sx = np.array([[0, 1], [1, 0]], dtype=complex) / 2
Sy = np.array([[0, -1j], [1j, 0]], dtype=complex) / 2
Sz = np.array([[1, 0], [0, -1]], dtype=complex) / 2
RHO0 = np.diag([1, 0]).astype(complex)
RHO1 = np.diag([0, 1]).astype(complex)


def test_pi_pulse_state_transfer():
    N, T = 10, 1.0
    x = np.zeros((2, N)); x[0] = np.pi / T
    F, G, P, S, C = orc.fom_and_gradient_grape(0 * Sz, [Sx, Sy], x, T, RHO0, RHO1, orc.STATE_TRANSFER, return_stores=True)
    assert np.allclose(S[-1], RHO1, atol=1e-14)
    assert abs(F - 0.75) < 1e-14            # C1(rho, rho) = 1 - |1/2|^2
    assert np.max(np.abs(G)) < 1e-14


def test_pi_pulse_unitary():
    N, T = 10, 1.0
    x = np.zeros((2, N)); x[0] = np.pi / T
    F, G = orc.fom_and_gradient_grape(0 * Sz, [Sx, Sy], x, T, np.eye(2, dtype=complex), 2 * Sx, orc.UNITARY_GATE)
    assert abs(F + 4.0) < 1e-13             # tau = +2i, Re(tau^2) = -4
    assert np.max(np.abs(G)) < 1e-13
    Fs, Gs = orc.fom_and_gradient_grape(Sz, [Sx, Sy], np.random.default_rng(0).random((2, 10)), T,
                                        np.eye(2, dtype=complex), 2 * Sx, orc.UNITARY_GATE, orc.REF_STATIC)
    Fi, Gi = orc.fom_and_gradient_grape(Sz, [Sx, Sy], np.random.default_rng(0).random((2, 10)), T,
                                        np.eye(2, dtype=complex), 2 * Sx, orc.UNITARY_GATE, orc.REF_INPLACE)
    assert Fs == Fi and np.allclose(Gs, -Gi)   # GRAPE.jl:272 vs :290 differ by sign only


def test_zero_pulse_diagonal_drift_closed_form():
    D, N, T = 4, 6, 0.7
    d = np.array([0.3, -1.1, 0.5, 2.0])
    A = np.diag(d).astype(complex)
    P = orc.pw_prop_save(A, [np.zeros((D, D), dtype=complex)], np.zeros((1, N)), T / N)
    for p in P:
        assert np.allclose(p, np.diag(np.exp(-1j * T / N * d)), atol=1e-15)
    U = orc.pw_evolve(A, [np.zeros((D, D), dtype=complex)], np.zeros((1, N)), T / N, np.eye(D, dtype=complex))
    assert np.allclose(U, np.diag(np.exp(-1j * T * d)), atol=1e-14)


def test_expm_against_mpmath():
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 50
    for D, seed in ((2, 1), (4, 2), (8, 3)):
        A, B, _, _ = random_system(D, 2, seed)
        x = np.random.default_rng(seed).uniform(-1, 1, (2, 1))
        dt = 0.05
        P = orc.pw_prop_save(A, B, x, dt)[0]
        H = orc.pw_ham(A, B, x, 0)
        M = mp.matrix(D, D)
        for i in range(D):
            for j in range(D):
                M[i, j] = mp.mpc(-1j * dt * H[i, j])
        E = mp.expm(M)
        ref = np.array([[complex(E[i, j]) for j in range(D)] for i in range(D)])
        assert np.max(np.abs(P - ref)) < 5e-16


@pytest.mark.parametrize("sys_type", [orc.STATE_TRANSFER, orc.UNITARY_GATE, orc.COHERENCE_TRANSFER])
def test_exact_gradient_matches_finite_differences(sys_type):
    D, K, N, T = 4, 3, 8, 1.2
    A, B, Xi, Xt = random_system(D, K, seed=5, hermitian=(sys_type != orc.COHERENCE_TRANSFER),
                                 unitary_targets=(sys_type == orc.UNITARY_GATE))
    x = np.random.default_rng(1).uniform(-1, 1, (K, N))
    F, G = orc.exact_fom_and_gradient(A, B, x, T, Xi, Xt, sys_type)
    assert abs(F - orc.exact_functional(A, B, x, T, Xi, Xt, sys_type)) < 1e-13
    h = 1e-6
    for (c, t) in [(0, 0), (1, 3), (2, 7), (0, 5)]:
        xp, xm = x.copy(), x.copy()
        xp[c, t] += h; xm[c, t] -= h
        fd = (orc.exact_functional(A, B, xp, T, Xi, Xt, sys_type) - orc.exact_functional(A, B, xm, T, Xi, Xt, sys_type)) / (2 * h)
        assert abs(fd - G[c, t]) < 1e-8 * max(1.0, np.max(np.abs(G)))


def test_overlap_is_slice_independent():
    """tr(S_t' C_t) (unitary) and tr(C_t' S_t) (density) do not depend on t."""
    A, B, Xi, Xt = random_system(4, 2, seed=8, unitary_targets=True)
    x = np.random.default_rng(4).uniform(-1, 1, (2, 9))
    for st in (orc.UNITARY_GATE, orc.STATE_TRANSFER):
        _, _, P, S, C = orc.fom_and_gradient_grape(A, B, x, 1.0, Xi, Xt, st, return_stores=True)
        taus = [np.trace(orc.dag(C[t]) @ S[t]) for t in range(len(S))]
        assert np.max(np.abs(np.array(taus) - taus[0])) < 1e-12
        for p in P:
            assert np.max(np.abs(orc.dag(p) @ p - np.eye(4))) < 1e-13    # unitarity for Hermitian H


def test_first_order_gradient_converges_to_exact_as_dt_to_zero():
    """Density types: g_ref ~ d/du [-Re tr(Xt' rho_N)] to first order in dt."""
    A, B, Xi, Xt = random_system(2, 2, seed=3)
    errs = []
    for N in (20, 40, 80):
        x = np.tile(np.array([[0.3], [-0.2]]), (1, N))
        _, g = orc.fom_and_gradient_grape(A, B, x, 1.0, Xi, Xt, orc.STATE_TRANSFER)
        h = 1e-6
        xp, xm = x.copy(), x.copy(); xp[0, N // 2] += h; xm[0, N // 2] -= h

        def f(xx):
            U = orc.pw_evolve(A, B, xx, 1.0 / N, np.eye(2, dtype=complex))
            return -np.real(np.trace(orc.dag(Xt) @ U @ Xi @ orc.dag(U)))
        fd = (f(xp) - f(xm)) / (2 * h)
        errs.append(abs(g[0, N // 2] - fd) / abs(fd))
    assert errs[1] < 0.62 * errs[0] and errs[2] < 0.62 * errs[1]


def test_ensemble_weighting():
    members = [random_system(2, 2, seed=20 + k) for k in range(3)]
    x = np.random.default_rng(6).random((2, 5))
    wts = [0.2, 0.5, 0.3]
    F, G = orc.ensemble_fom_and_gradient(members, wts, x, 1.0, orc.STATE_TRANSFER)
    Fs = Gs = 0
    for w, (A, B, Xi, Xt) in zip(wts, members):
        f, g = orc.fom_and_gradient_grape(A, B, x, 1.0, Xi, Xt, orc.STATE_TRANSFER)
        Fs, Gs = Fs + w * f, Gs + w * g
    assert abs(F - Fs) < 1e-15 and np.max(np.abs(G - Gs)) < 1e-15


def test_julia_exp_rung_for_baseline_configs():
    """All five BASELINE configs sit on the Pade-5 rung (E = 3 products) of Julia's exp!."""
    import quoptimalcontrol_jl_b200 as qoc
    for cfg in (qoc.configs.config1(), qoc.configs.config2(N=50), qoc.configs.config3(N=50),
                qoc.configs.config4(N=20, grid=2), qoc.configs.config5(N=4, n=8)):
        A, B, _, _ = cfg["members"][0]
        dt = cfg["T"] / cfg["N"]
        for i in range(cfg["N"]):
            deg, prods, s = orc.julia_exp_rung(-1j * dt * orc.pw_ham(A, B, cfg["x"], i))
            assert (deg, prods, s) == (5, 3, 0), (cfg["name"], i, deg)


# ---- algebraic identities the CUDA paths rely on, pinned against the line-by-line oracle on the CPU -------------------
def _closed_system(D, K, N, seed):
    rng = np.random.default_rng(seed)
    herm = lambda: (lambda Z: (Z + Z.conj().T) / 2)(rng.normal(size=(D, D)) + 1j * rng.normal(size=(D, D)))
    return herm(), [herm() for _ in range(K)], rng.uniform(-1, 1, (K, N)), rng


def test_closed_system_conjugation_recursion():
    """Hermitian drift and controls: W_t = S_t C_t' - C_t' S_t = U_t W_0 U_t' with U_t = P_{t-1} ... P_0, so one forward
    recursion replaces both sweeps; the gradient from it equals the oracle's."""
    D, K, N, T = 5, 2, 9, 1.3
    A, B, x, rng = _closed_system(D, K, N, 1)
    rho = lambda: (lambda Z: Z @ Z.conj().T)(rng.normal(size=(D, D)) + 1j * rng.normal(size=(D, D)))
    Xi, Xt = rho(), rho()
    dt = T / N
    P = orc.pw_prop_save(A, B, x, dt)
    UN = np.eye(D, dtype=complex)
    for p in P:
        UN = p @ UN
    C0 = UN.conj().T @ Xt @ UN
    W = Xi @ C0.conj().T - C0.conj().T @ Xi
    G = np.zeros((K, N))
    for t in range(N):
        for c in range(K):
            G[c, t] = np.real(1j * dt * np.trace(B[c] @ W))
        W = P[t] @ W @ P[t].conj().T
    Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, orc.STATE_TRANSFER)
    assert np.max(np.abs(G - Go)) < 1e-12 * max(1.0, np.max(np.abs(Go)))
    tau = np.trace(C0.conj().T @ Xi)
    assert abs((1 - abs(tau / D) ** 2) - Fo) < 1e-12 * max(1.0, abs(Fo))


def test_pure_state_vector_identities():
    """Xi = psi psi', Xt = phi phi' on a closed system: g[c,t] = -2 dt Im(conj(o_t) chi_t' B_c psi_t), o_t = chi_t' psi_t,
    fom = 1 - |o|^4 / D^2 (the formulas of csrc/pure_state.cuh) reproduce the oracle."""
    D, K, N, T = 6, 3, 8, 0.9
    A, B, x, rng = _closed_system(D, K, N, 2)
    ket = lambda: (lambda v: v / np.linalg.norm(v))(rng.normal(size=D) + 1j * rng.normal(size=D))
    psi0, phi = ket(), ket()
    Xi, Xt = np.outer(psi0, psi0.conj()), np.outer(phi, phi.conj())
    dt = T / N
    P = orc.pw_prop_save(A, B, x, dt)
    psi = [psi0]
    for p in P:
        psi.append(p @ psi[-1])
    chi = [None] * (N + 1)
    chi[N] = phi
    for t in range(N - 1, -1, -1):
        chi[t] = P[t].conj().T @ chi[t + 1]
    G = np.zeros((K, N))
    for t in range(N):
        o = np.vdot(chi[t], psi[t])
        for c in range(K):
            G[c, t] = -2 * dt * np.imag(np.conj(o) * np.vdot(chi[t], B[c] @ psi[t]))
    Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, orc.STATE_TRANSFER)
    assert np.max(np.abs(G - Go)) < 1e-13 * max(1.0, np.max(np.abs(Go)))
    o = np.vdot(chi[N - 1], psi[N - 1])
    assert abs((1 - abs(o) ** 4 / D ** 2) - Fo) < 1e-13


def test_slice_range_boundary_operators():
    """Slice-parallel algebra: a range evaluated with S_lo = L Xi L', C_hi = R' Xt R as its initial / target
    operators returns the full figure of merit and exactly its slices' gradient entries."""
    D, K, N, T, cut = 4, 2, 10, 1.1, 4
    A, B, x, rng = _closed_system(D, K, N, 3)
    A = A + 0.05j * rng.normal(size=(D, D))                      # not Hermitian: the identity does not need unitarity
    rho = lambda: (lambda Z: Z @ Z.conj().T)(rng.normal(size=(D, D)) + 1j * rng.normal(size=(D, D)))
    Xi, Xt = rho(), rho()
    I = np.eye(D, dtype=complex)
    dt = T / N
    U0, U1 = orc.pw_evolve(A, B, x[:, :cut], dt, I), orc.pw_evolve(A, B, x[:, cut:], dt, I)
    Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, orc.COHERENCE_TRANSFER)
    F0, G0 = orc.fom_and_gradient_grape(A, B, x[:, :cut], dt * cut, Xi, U1.conj().T @ Xt @ U1, orc.COHERENCE_TRANSFER)
    F1, G1 = orc.fom_and_gradient_grape(A, B, x[:, cut:], dt * (N - cut), U0 @ Xi @ U0.conj().T, Xt, orc.COHERENCE_TRANSFER)
    assert abs(F0 - Fo) < 1e-12 * max(1.0, abs(Fo)) and abs(F1 - Fo) < 1e-12 * max(1.0, abs(Fo))
    assert np.max(np.abs(np.concatenate([G0, G1], axis=1) - Go)) < 1e-12 * max(1.0, np.max(np.abs(Go)))
