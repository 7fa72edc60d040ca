"""CPU oracle for the GRAPE fidelity+gradient hot path of QuOptimalControl.jl.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (`quoptimalcontrol.jl_b200/`) may import this
module; only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg do.

PARITY UNPINNED.  The reference ships no golden vectors, known-answer tests or fixtures for this path
(its 12 test sets only assert one-sided bounds on an optimiser's final minimum from unseeded random
starts), and Julia is not installed here, so the reference itself cannot be executed.  This oracle is a
line-by-line restatement of the reference's formulas (citations below, all relative to
/root/reference/) and is pinned instead against (i) analytic known answers, (ii) 50-digit mpmath
matrix exponentials, (iii) finite differences of the reference's own AD functional
(tests/test_oracle.py), and (iv) the assertions of the reference's own 12 test sets, replayed on this
oracle with seeded guesses (tests/test_oracle_reference_scenarios.py): it meets every bound the reference
checks.  Value-level parity with a running Julia reference stays unpinned.

The matrix exponential is the one piece of third-party arithmetic on the path: the reference calls
`LinearAlgebra.exp(::Matrix{ComplexF64})` (Julia stdlib, version = the user's Julia >= 1.6;
Higham-2005 scaling-and-squaring Pade) at src/timeevolution.jl:36,53,108.  `scipy.linalg.expm`
(Al-Mohy/Higham 2009 Pade, same algorithm family, ~1e-15 relative) stands in for it here;
`julia_exp_rung` restates the stdlib degree ladder only to *report* which Pade rung (how many matrix
products) the reference would execute for a given input.

Conventions: all matrices complex128; the pulse `x` has shape (K, N) like the reference's
`control_array[j, i]` (control j, slice i).  0-based indices here, 1-based in the reference.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm as _expm

STATE_TRANSFER = 0      # src/problems.jl:8
UNITARY_GATE = 1        # src/problems.jl:9
COHERENCE_TRANSFER = 2  # src/problems.jl:10

REF_INPLACE = 0   # grad_func! (src/GRAPE.jl:261-287)
REF_STATIC = 1    # grad_func  (src/GRAPE.jl:289-303)


def dag(X):
    return X.conj().T


# --------------------------------------------------------------------------- time evolution
def pw_ham(A, B, x, i):
    """H_i = A + sum_j B[j]*x[j,i], accumulated from zero with j ascending and A added last
    (src/timeevolution.jl:103-108)."""
    H = np.zeros_like(A, dtype=np.complex128)
    for j in range(len(B)):
        H = H + B[j] * x[j, i]
    return H + A


def pw_prop_save(A, B, x, dt):
    """P_i = exp(-1im*dt*(Htot + A)) for every slice (src/timeevolution.jl:98-110)."""
    N = x.shape[1]
    return [_expm((-1.0j * dt) * pw_ham(A, B, x, i)) for i in range(N)]


def pw_ham_save(A, B, x):
    """src/timeevolution.jl:64-75."""
    return [pw_ham(A, B, x, i) for i in range(x.shape[1])]


def pw_gen_save(A, B, x, duration):
    """src/timeevolution.jl:80-92."""
    N = x.shape[1]
    dt = duration / N
    return [pw_ham(A, B, x, i) * (-1.0j * dt) for i in range(N)]


def pw_evolve(A, B, x, dt, U0):
    """U = P_N ... P_1 U0 (src/timeevolution.jl:28-39).  Htot starts from A here."""
    U = U0
    for i in range(x.shape[1]):
        H = A
        for j in range(len(B)):
            H = H + B[j] * x[j, i]
        U = _expm((-1.0j * dt) * H) @ U
    return U


# --------------------------------------------------------------------------- cost functions
def C1(KT, KN):
    """1 - |tr(KT' KN)/D|^2 (src/cost_functions.jl:13-17)."""
    D = KT.shape[0]
    return 1.0 - abs(np.trace(dag(KT) @ KN) / D) ** 2


def commutator(A, B):
    """src/tools.jl:17-19."""
    return A @ B - B @ A


def fom_func(sys_type, t, S, C):
    """5-argument fom_func (src/cost_functions.jl:99-111)."""
    if sys_type == UNITARY_GATE:
        tau = np.trace(dag(S[t]) @ C[t])
        return float(np.real(tau * tau))
    return float(np.real(C1(C[t], S[t])))


# --------------------------------------------------------------------------- GRAPE evaluator
def fom_and_gradient_grape(A, B, x, T, Xi, Xt, sys_type, variant=REF_INPLACE, return_stores=False):
    """_fom_and_gradient_GRAPE! (src/GRAPE.jl:25-96) / _fom_and_gradient_sGRAPE (:103-166).

    Returns (fom, grad[K, N]).  Reference loop order and operation count are kept: the gradient
    double loop does its three GEMMs per (c, t)."""
    K, N = x.shape
    dt = T / N                                                # GRAPE.jl:42
    S = [None] * (N + 1)
    C = [None] * (N + 1)
    S[0] = np.array(Xi, dtype=np.complex128)                  # GRAPE.jl:44
    C[N] = np.array(Xt, dtype=np.complex128)                  # GRAPE.jl:45
    P = pw_prop_save(A, B, x, dt)                             # GRAPE.jl:49
    unitary = sys_type == UNITARY_GATE
    for t in range(N):                                        # GRAPE.jl:53-63
        if unitary:
            S[t + 1] = P[t] @ S[t]                            # GRAPE.jl:226
        else:
            store = S[t] @ dag(P[t])                          # GRAPE.jl:245
            S[t + 1] = P[t] @ store                           # GRAPE.jl:246
    for t in reversed(range(N)):                              # GRAPE.jl:65-75
        if unitary:
            C[t] = dag(P[t]) @ C[t + 1]                       # GRAPE.jl:228
        else:
            store = C[t + 1] @ P[t]                           # GRAPE.jl:248
            C[t] = dag(P[t]) @ store                          # GRAPE.jl:249
    g = np.zeros((K, N))
    for c in range(K):                                        # GRAPE.jl:79-92
        for t in range(N):
            if unitary:
                if variant == REF_INPLACE:                    # GRAPE.jl:271-272
                    store = dag(S[t]) @ C[t]
                    g[c, t] = 2.0 * np.real((1.0j * dt) * np.trace(dag(C[t]) @ B[c] @ S[t]) * np.trace(store))
                else:                                         # GRAPE.jl:290
                    g[c, t] = 2.0 * np.real((-1.0j * dt) * np.trace(dag(C[t]) @ B[c] @ S[t])
                                            * np.trace(dag(S[t]) @ C[t]))
            else:                                             # GRAPE.jl:285-286, :302
                store = dag(C[t]) @ commutator(B[c], S[t])
                g[c, t] = np.real(np.trace((1.0j * dt) * store))
    fom = fom_func(sys_type, N - 1, S, C)                     # GRAPE.jl:77,94 (1-based t = N)
    if return_stores:
        return fom, g, P, S, C
    return fom, g


def ensemble_fom_and_gradient(members, wts, x, T, sys_type, variant=REF_INPLACE):
    """Ensemble closure of solve(::EnsembleProblem, ::GRAPE) (src/solve.jl:164-196).
    `members` is a list of (A, B, Xi, Xt) as produced by init_ensemble (src/tools.jl:42-53)."""
    F = 0.0
    K, N = x.shape
    grads = np.zeros((len(members), K, N))
    for k, (A, B, Xi, Xt) in enumerate(members):              # solve.jl:166, serial, k ascending
        fk, gk = fom_and_gradient_grape(A, B, x, T, Xi, Xt, sys_type, variant)
        F += fk * wts[k]                                      # solve.jl:171-186
        grads[k] = gk
    G = np.sum(grads * np.asarray(wts)[:, None, None], axis=0)  # solve.jl:191
    return F, G


# --------------------------------------------------------------------------- exact (ADGRAPE) semantics
def exact_functional(A, B, x, T, Xi, Xt, sys_type):
    """_get_functional (src/solve.jl:268-290): C1(Xt, U Xi U') or C1(Xt, U Xi), U from pw_evolve with
    U0 = I.  The reference has no CoherenceTransfer method; the density sandwich is used for it."""
    K, N = x.shape
    D = A.shape[0]
    U = pw_evolve(A, B, x, T / N, np.eye(D, dtype=np.complex128))
    ev = U @ Xi if sys_type == UNITARY_GATE else U @ Xi @ dag(U)
    return float(C1(Xt, ev))


def exact_fom_and_gradient(A, B, x, T, Xi, Xt, sys_type):
    """F = exact_functional, G = dF/dx exactly (what real(Zygote.gradient(functional)) returns,
    src/GRAPE.jl:14-18), by the augmented-matrix (Van Loan) derivative of every slice propagator:
    dP_{c,t} = upper-right block of exp([[G_t, -i dt B_c], [0, G_t]])."""
    K, N = x.shape
    D = A.shape[0]
    dt = T / N
    P = pw_prop_save(A, B, x, dt)
    unitary = sys_type == UNITARY_GATE
    S = [None] * (N + 1)
    C = [None] * (N + 1)
    S[0] = np.array(Xi, dtype=np.complex128)
    C[N] = np.array(Xt, dtype=np.complex128)
    for t in range(N):
        S[t + 1] = P[t] @ S[t] if unitary else P[t] @ S[t] @ dag(P[t])
    for t in reversed(range(N)):
        C[t] = dag(P[t]) @ C[t + 1] if unitary else dag(P[t]) @ C[t + 1] @ P[t]
    tau = np.trace(dag(Xt) @ S[N])
    F = 1.0 - abs(tau / D) ** 2
    G = np.zeros((K, N))
    Z = np.zeros((D, D), dtype=np.complex128)
    for t in range(N):
        Gt = (-1.0j * dt) * pw_ham(A, B, x, t)
        for c in range(K):
            aug = np.block([[Gt, (-1.0j * dt) * B[c]], [Z, Gt]])
            dP = _expm(aug)[:D, D:]
            if unitary:
                dtau = np.trace(dag(C[t + 1]) @ dP @ S[t])
            else:
                dtau = np.trace(dag(C[t + 1]) @ dP @ S[t] @ dag(P[t])) \
                    + np.trace(dag(C[t + 1]) @ P[t] @ S[t] @ dag(dP))
            G[c, t] = -(2.0 / D ** 2) * np.real(np.conj(tau) * dtau)
    return float(F), G


def ensemble_exact(members, wts, x, T, sys_type):
    """_get_ensemble_functional (src/solve.jl:312-361): sum_k wts[k]*f_k and its gradient."""
    F = 0.0
    G = np.zeros(x.shape)
    for k, (A, B, Xi, Xt) in enumerate(members):
        fk, gk = exact_fom_and_gradient(A, B, x, T, Xi, Xt, sys_type)
        F += fk * wts[k]
        G += gk * wts[k]
    return F, G


# --------------------------------------------------------------------------- which Pade rung would Julia's exp! take
def julia_exp_rung(M):
    """Restates the degree ladder of Julia's LinearAlgebra.exp! (stdlib dense.jl; Higham 2005), from
    the published algorithm: returns (pade_degree, n_products, n_squarings) for the 1-norm of M.
    Balancing (gebal) is ignored; it can only lower the norm."""
    nA = np.linalg.norm(M, 1)
    if nA <= 2.1:
        for theta, deg, prods in ((0.015, 3, 2), (0.25, 5, 3), (0.95, 7, 4), (2.1, 9, 5)):
            if nA <= theta:
                return deg, prods, 0
    s = max(0, int(np.ceil(np.log2(nA / 5.4))))
    return 13, 6, s
