"""GPU parity of the large-dimension path (D > 16: tiled DMMA GEMM pipeline) against the numpy oracle."""
import numpy as np
import pytest

import quoptimalcontrol_jl_b200 as qoc
from oracle import grape_oracle as orc
from conftest import assert_parity, random_system

pytestmark = pytest.mark.gpu
SYS = {"state": orc.STATE_TRANSFER, "unitary": orc.UNITARY_GATE, "coherence": orc.COHERENCE_TRANSFER}


@pytest.mark.parametrize("D,N", [(64, 12), (20, 7), (32, 9), (90, 5), (100, 5), (128, 9), (64, 1), (64, 2), (64, 3)])
@pytest.mark.parametrize("sys_name", ["state", "unitary", "coherence"])
def test_first_order_dense_random(D, N, sys_name):
    K, T = 3, 0.8
    A, B, Xi, Xt = random_system(D, K, seed=600 + D + N, hermitian=(sys_name != "coherence"),
                                 unitary_targets=(sys_name == "unitary"))
    x = np.random.default_rng(D + N).uniform(-1, 1, (K, N))
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], T, N, SYS[sys_name]) as ev:
        F, G = ev.eval(x)
        F0, _ = ev.eval(x, want_grad=False)
        assert ev.stats()["path"] == 2
    Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, SYS[sys_name])
    assert_parity(F, G, Fo, Go)
    assert_parity(F0, None, Fo, None)


def test_static_sign_and_ensemble():
    D, K, N, T, M = 64, 2, 6, 0.5, 3
    members = [random_system(D, K, seed=700 + k, unitary_targets=True) for k in range(M)]
    wts = [0.5, 0.2, 0.3]
    x = np.random.default_rng(5).uniform(-1, 1, (K, N))
    with qoc.GrapeEvaluator(members, T, N, orc.UNITARY_GATE, wts=wts, convention="static") as ev:
        F, G = ev.eval(x)
    Fo, Go = orc.ensemble_fom_and_gradient(members, wts, x, T, orc.UNITARY_GATE, orc.REF_STATIC)
    assert_parity(F, G, Fo, Go)


@pytest.mark.parametrize("batch", [None, 1, 4])
@pytest.mark.parametrize("sys_name", ["state", "unitary"])
@pytest.mark.parametrize("D,N", [(32, 11), (20, 7), (64, 5)])
def test_chain_batching(monkeypatch, batch, sys_name, D, N):
    """Closed-system chains (members x pulses) ride in the GEMM batch dimension, QOC_BIG_BATCH at a time: an
    M = 3, R = 2 ensemble evaluated in batches of 6 (default), 1 and 4 + 2 (ragged last batch) agrees with the oracle."""
    if batch is None:
        monkeypatch.delenv("QOC_BIG_BATCH", raising=False)
    else:
        monkeypatch.setenv("QOC_BIG_BATCH", str(batch))
    K, T, M, R = 2, 0.7, 3, 2
    members = [random_system(D, K, seed=900 + 7 * k + D, unitary_targets=(sys_name == "unitary")) for k in range(M)]
    wts = [0.5, 0.2, 0.3]
    xs = np.random.default_rng(N + D).uniform(-1, 1, (R, K, N))
    with qoc.GrapeEvaluator(members, T, N, SYS[sys_name], wts=wts, n_pulses=R) as ev:
        F, G = ev.eval(xs)
        F2, G2 = ev.eval(xs)          # cached batch descriptor on the second call
    for r in range(R):
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, xs[r], T, SYS[sys_name])
        assert_parity(F[r], G[r], Fo, Go)
        assert_parity(F2[r], G2[r], Fo, Go)


def test_large_norm_squarings():
    D, K, N, T = 64, 2, 4, 2.0
    A, B, Xi, Xt = random_system(D, K, seed=9, scale=6.0)
    x = np.random.default_rng(1).uniform(-1, 1, (K, N))
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], T, N, orc.STATE_TRANSFER) as ev:
        F, G = ev.eval(x)
    Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, orc.STATE_TRANSFER)
    assert_parity(F, G, Fo, Go, ftol=1e-9, gtol=1e-7)


@pytest.mark.parametrize("pure_state", [False, True])
@pytest.mark.parametrize("n,N", [(6, 40), (8, 6)])
def test_config5_reduced(monkeypatch, n, N, pure_state):
    monkeypatch.setenv("QOC_PURE_STATE", "1" if pure_state else "0")      # "1": also below the size where the library picks it
    cfg = qoc.configs.config5(N=N, n=n)
    A, B, Xi, Xt = cfg["members"][0]
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], N, cfg["sys_type"], pure_state=pure_state) as ev:
        F, G = ev.eval(cfg["x"])
        assert ev.stats()["path"] == (3 if pure_state else 2)
    Fo, Go = orc.fom_and_gradient_grape(A, B, cfg["x"], cfg["T"], Xi, Xt, cfg["sys_type"])
    assert_parity(F, G, Fo, Go)


def test_propagators_and_total_big():
    D, K, N, T = 64, 2, 9, 1.1
    A, B, _, _ = random_system(D, K, seed=31)
    x = np.random.default_rng(2).uniform(-1, 1, (K, N))
    dt = T / N
    P = qoc.pw_prop_save(A, B, x, K, N, dt)
    for a, b in zip(P, orc.pw_prop_save(A, B, x, dt)):
        assert np.max(np.abs(a - b)) < 1e-13
    for a, b in zip(qoc.pw_ham_save(A, B, x, K, N), orc.pw_ham_save(A, B, x)):
        assert np.max(np.abs(a - b)) < 1e-14
    I = np.eye(D, dtype=complex)
    assert np.max(np.abs(qoc.pw_evolve(A, B, x, K, dt, N, I) - orc.pw_evolve(A, B, x, dt, I))) < 1e-12


@pytest.mark.parametrize("D,N", [(64, 6), (40, 4), (24, 5)])
@pytest.mark.parametrize("sys_name", ["state", "unitary", "coherence"])
def test_exact_gradient_big(D, N, sys_name):
    """Exact (ADGRAPE-semantics) gradient on the tiled-GEMM path: Frechet derivative of the Taylor-8 scheme."""
    K, T = 2, 0.6
    A, B, Xi, Xt = random_system(D, K, seed=800 + D, hermitian=(sys_name != "coherence"),
                                 unitary_targets=(sys_name == "unitary"))
    x = np.random.default_rng(D + 1).uniform(-1, 1, (K, N))
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], T, N, SYS[sys_name], gradient="exact") as ev:
        F, G = ev.eval(x)
        F0, _ = ev.eval(x, want_grad=False)
    Fo, Go = orc.exact_fom_and_gradient(A, B, x, T, Xi, Xt, SYS[sys_name])
    assert_parity(F, G, Fo, Go)
    assert_parity(F0, None, Fo, None)


@pytest.mark.parametrize("scale", [1.5, 3.0])
def test_exact_gradient_big_with_squarings(scale):
    D, K, N, T = 64, 2, 3, 1.0
    A, B, Xi, Xt = random_system(D, K, seed=12, scale=scale)
    x = np.random.default_rng(4).uniform(-1, 1, (K, N))
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], T, N, orc.STATE_TRANSFER, gradient="exact") as ev:
        F, G = ev.eval(x)
    Fo, Go = orc.exact_fom_and_gradient(A, B, x, T, Xi, Xt, orc.STATE_TRANSFER)
    assert_parity(F, G, Fo, Go, ftol=1e-9, gtol=1e-7)


def test_exact_too_many_squarings_is_loud():
    A, B, Xi, Xt = random_system(64, 1, seed=1, scale=4000.0)
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], 1.0, 2, orc.STATE_TRANSFER, gradient="exact") as ev:
        with pytest.raises(qoc.QocError) as e:
            ev.eval(np.ones((1, 2)))
    assert e.value.status == qoc._lib.QOC_EUNSUPPORTED


# ------------------------------------------------------------------------------------------------------------------
# pure-state vector path (path 3): StateTransfer between rank-1 states on a sparse closed system
def _spin_chain(n, seed, complex_states=True):
    """n-qubit Ising chain with local x/y controls and random (complex) product-free pure states."""
    cfg = qoc.configs.config5(N=4, n=n)
    A, B, _, _ = cfg["members"][0]
    rng = np.random.default_rng(seed)
    D = 2 ** n
    psi = rng.normal(size=D) + (1j * rng.normal(size=D) if complex_states else 0)
    phi = rng.normal(size=D) + (1j * rng.normal(size=D) if complex_states else 0)
    psi /= np.linalg.norm(psi); phi /= np.linalg.norm(phi)
    return A, B, np.outer(psi, psi.conj()), np.outer(phi, phi.conj())


@pytest.mark.parametrize("n,N,T", [(5, 9, 1.0), (6, 17, 2.5), (7, 5, 0.7), (5, 6, 60.0)])
def test_pure_state_path_matches_oracle(monkeypatch, n, N, T):
    """T = 60 with 6 slices drives ||dt H|| to ~40: exercises the Taylor sub-stepping."""
    monkeypatch.setenv("QOC_PURE_STATE", "1")
    A, B, Xi, Xt = _spin_chain(n, seed=40 + n)
    K = len(B)
    x = np.random.default_rng(n + N).uniform(-1, 1, (K, N))
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], T, N, orc.STATE_TRANSFER) as ev:
        F, G = ev.eval(x)
        F0, _ = ev.eval(x, want_grad=False)
        assert ev.stats()["path"] == 3
    Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, orc.STATE_TRANSFER)
    assert_parity(F, G, Fo, Go)
    assert_parity(F0, None, Fo, None)
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], T, N, orc.STATE_TRANSFER, pure_state=False) as ev:     # the flag wins over the env
        Fd, Gd = ev.eval(x)
        assert ev.stats()["path"] == 2
    assert_parity(F, G, Fd, Gd)


def test_pure_state_ensemble_and_pulses(monkeypatch):
    monkeypatch.setenv("QOC_PURE_STATE", "1")
    n, N, T, M, R = 5, 8, 1.3, 3, 2
    A, B, Xi, Xt = _spin_chain(n, seed=77)
    members = []
    for k in range(M):     # members differ in drift, control scale and states
        _, _, Xik, Xtk = _spin_chain(n, seed=78 + k)
        members.append((A * (1 + 0.1 * k), [b * (1 - 0.05 * k) for b in B], Xik, Xtk))
    wts = [0.2, 0.5, 0.3]
    xs = np.random.default_rng(3).uniform(-1, 1, (R, len(B), N))
    with qoc.GrapeEvaluator(members, T, N, orc.STATE_TRANSFER, wts=wts, n_pulses=R) as ev:
        F, G = ev.eval(xs)
        assert ev.stats()["path"] == 3
    for r in range(R):
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, xs[r], T, orc.STATE_TRANSFER)
        assert_parity(F[r], G[r], Fo, Go)


def test_pure_state_path_not_taken_when_it_does_not_apply(monkeypatch):
    monkeypatch.setenv("QOC_PURE_STATE", "1")
    n, N, T = 5, 4, 1.0
    A, B, Xi, Xt = _spin_chain(n, seed=5)
    x = np.random.default_rng(0).uniform(-1, 1, (len(B), N))
    mixed = 0.5 * Xi + 0.5 * Xt                                  # rank 2
    for members, sys_type in [([(A, B, mixed, Xt)], orc.STATE_TRANSFER),                 # mixed initial state
                              ([(A, B, Xi, Xt)], orc.COHERENCE_TRANSFER),                # other problem type
                              ([random_system(32, 2, seed=1)], orc.STATE_TRANSFER)]:     # dense system
        xk = np.random.default_rng(0).uniform(-1, 1, (len(members[0][1]), N))
        with qoc.GrapeEvaluator(members, T, N, sys_type) as ev:
            F, G = ev.eval(xk)
            assert ev.stats()["path"] == 2
        Am, Bm, Xim, Xtm = members[0]
        Fo, Go = orc.fom_and_gradient_grape(Am, Bm, xk, T, Xim, Xtm, sys_type)
        assert_parity(F, G, Fo, Go)


@pytest.mark.parametrize("D,K,path", [(24, 1, 3), (40, 3, 3), (36, 2, 3), (17, 1, 2), (24, 2, 2)])
def test_pure_state_path_general_sparse_system(monkeypatch, D, K, path):
    """Not a qubit register: banded complex Hermitian drift and controls (a few diagonals), D not a multiple of 32.
    The union pattern has 5 (K = 1) or 7 entries per row; with more than D/4 the dense path is kept (path 2)."""
    monkeypatch.setenv("QOC_PURE_STATE", "1")
    rng = np.random.default_rng(D)
    def banded(offsets):
        Hm = np.zeros((D, D), dtype=complex)
        for o in offsets:
            v = rng.normal(size=D - o) + (1j * rng.normal(size=D - o) if o else 0)
            Hm += np.diag(v, o) + (np.diag(v.conj(), -o) if o else 0)
        return Hm
    A = banded([0, 1])
    B = [banded([1, 3]) if j % 2 == 0 else banded([2]) for j in range(K)]
    psi = rng.normal(size=D) + 1j * rng.normal(size=D); psi /= np.linalg.norm(psi)
    phi = rng.normal(size=D) + 1j * rng.normal(size=D); phi /= np.linalg.norm(phi)
    Xi, Xt = np.outer(psi, psi.conj()), np.outer(phi, phi.conj())
    N, T = 11, 0.9
    x = rng.uniform(-1, 1, (K, N))
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], T, N, orc.STATE_TRANSFER) as ev:
        F, G = ev.eval(x)
        assert ev.stats()["path"] == path
    Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, orc.STATE_TRANSFER)
    assert_parity(F, G, Fo, Go)


def test_pure_state_path_selection_by_size():
    """Without the override the library keeps single small chains on the dense path and takes the vector path from D = 128,
    from 3 chains at D = 64 and from 16 chains below."""
    for n, M, want in [(5, 1, 2), (5, 16, 3), (6, 2, 2), (6, 3, 3), (7, 1, 3)]:
        A, B, Xi, Xt = _spin_chain(n, seed=n)
        members = [(A * (1 + 0.01 * k), B, Xi, Xt) for k in range(M)]
        x = np.random.default_rng(n).uniform(-1, 1, (len(B), 3))
        with qoc.GrapeEvaluator(members, 0.5, 3, orc.STATE_TRANSFER) as ev:
            F, G = ev.eval(x)
            assert ev.stats()["path"] == want, (n, M)
        Fo, Go = orc.ensemble_fom_and_gradient(members, np.ones(M), x, 0.5, orc.STATE_TRANSFER)
        assert_parity(F, G, Fo, Go)


def test_pure_state_path_without_controls(monkeypatch):
    monkeypatch.setenv("QOC_PURE_STATE", "1")
    A, B, Xi, Xt = _spin_chain(5, seed=3)
    with qoc.GrapeEvaluator([(A, [], Xi, Xt)], 1.0, 5, orc.STATE_TRANSFER) as ev:
        F, G = ev.eval(np.zeros((0, 5)))
        path = ev.stats()["path"]
    Fo, _ = orc.fom_and_gradient_grape(A, [], np.zeros((0, 5)), 1.0, Xi, Xt, orc.STATE_TRANSFER)
    assert path in (2, 3) and abs(F - Fo) <= 1e-10 * max(1.0, abs(Fo))


# ------------------------------------------------------------------------------------------------------------------
# slice-range evaluation on one device: the C-ABI pieces of the slice-parallel multi-GPU mode
@pytest.mark.parametrize("sys_name", ["state", "unitary", "coherence"])
@pytest.mark.parametrize("gradient", ["first_order", "exact"])
def test_set_states_and_eval_continue(sys_name, gradient):
    """Two handles own the slice ranges [0, 5) and [5, 12) of one D = 40 problem (what two ranks would hold): range
    propagators -> boundary operators -> qoc_set_states -> qoc_eval_continue reproduces the full evaluation."""
    if gradient == "exact" and sys_name == "coherence":
        pytest.skip("the reference defines no exact functional for CoherenceTransfer")
    D, K, N, T, cut = 40, 2, 12, 1.1, 5
    sys_type = SYS[sys_name]
    A, B, Xi, Xt = random_system(D, K, seed=300 + len(sys_name), hermitian=(sys_name != "coherence"),
                                 unitary_targets=(sys_name == "unitary"))
    x = np.random.default_rng(11).uniform(-1, 1, (K, N))
    I = np.eye(D, dtype=complex)
    ranges = [(0, cut), (cut, N)]
    evs = [qoc.GrapeEvaluator([(A, B, I, I)], T * (hi - lo) / N, hi - lo, sys_type, gradient=gradient, pure_state=False)
           for lo, hi in ranges]
    try:
        U = [ev.total_propagator(x[:, lo:hi]) for ev, (lo, hi) in zip(evs, ranges)]
        dagger = lambda m: m.conj().T
        if sys_name == "unitary":
            states = [(Xi, dagger(U[1]) @ Xt), (U[0] @ Xi, Xt)]
        else:
            states = [(Xi, dagger(U[1]) @ Xt @ U[1]), (U[0] @ Xi @ dagger(U[0]), Xt)]
        Fs, Gs = [], []
        for ev, (lo, hi), (s_lo, c_hi) in zip(evs, ranges, states):
            ev.total_propagator(x[:, lo:hi])          # the call eval_continue continues from
            ev.set_states(s_lo, c_hi)
            F, G = ev.eval_continue()
            Fs.append(F); Gs.append(G)
        with pytest.raises(qoc._lib.QocError):
            evs[0].eval(x[:, :cut]); evs[0].eval_continue()        # an ordinary eval in between invalidates the propagators
    finally:
        for ev in evs:
            ev.close()
    if gradient == "exact":
        Fo, Go = orc.exact_fom_and_gradient(A, B, x, T, Xi, Xt, sys_type)
    else:
        Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, sys_type)
    for F in Fs:
        assert abs(F - Fo) <= 1e-10 * max(1.0, abs(Fo))
    assert_parity(Fs[0], np.concatenate(Gs, axis=1), Fo, Go)
