"""GPU parity: libqocgrape.so (through the C ABI / ctypes) against the CPU oracle on identical seeded inputs."""
import numpy as np
import pytest

import quoptimalcontrol_jl_b200 as qoc
from oracle import grape_oracle as orc
from conftest import assert_parity, random_system

pytestmark = pytest.mark.gpu

SYS = {"state": orc.STATE_TRANSFER, "unitary": orc.UNITARY_GATE, "coherence": orc.COHERENCE_TRANSFER}


def _run(members, wts, T, N, sys_type, x, gradient, convention="inplace", R=1):
    with qoc.GrapeEvaluator(members, T, N, sys_type, wts=wts, gradient=gradient, convention=convention,
                            n_pulses=R) as ev:
        F, G = ev.eval(x)
        F0, _ = ev.eval(x, want_grad=False)
    return F, G, F0


@pytest.mark.parametrize("D", [2, 3, 4, 6, 8, 11, 16])
@pytest.mark.parametrize("sys_name", ["state", "unitary", "coherence"])
@pytest.mark.parametrize("variant", ["inplace", "static"])
def test_first_order_single(D, sys_name, variant):
    K, N, T = 3, 9, 1.3
    A, B, Xi, Xt = random_system(D, K, seed=100 + D, hermitian=(sys_name != "coherence"),
                                 unitary_targets=(sys_name == "unitary"))
    x = np.random.default_rng(D).uniform(-1, 1, (K, N))
    F, G, F0 = _run([(A, B, Xi, Xt)], None, T, N, SYS[sys_name], x, "first_order", variant)
    Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, SYS[sys_name],
                                        orc.REF_INPLACE if variant == "inplace" else orc.REF_STATIC)
    assert_parity(F, G, Fo, Go)
    assert_parity(F0, None, Fo, None)


@pytest.mark.parametrize("D", [2, 4, 5, 8, 16])
@pytest.mark.parametrize("sys_name", ["state", "unitary", "coherence"])
def test_exact_single(D, sys_name):
    K, N, T = 2, 7, 0.9
    A, B, Xi, Xt = random_system(D, K, seed=200 + D, hermitian=(sys_name != "coherence"),
                                 unitary_targets=(sys_name == "unitary"))
    x = np.random.default_rng(D + 50).uniform(-1, 1, (K, N))
    F, G, F0 = _run([(A, B, Xi, Xt)], None, T, N, SYS[sys_name], x, "exact")
    Fo, Go = orc.exact_fom_and_gradient(A, B, x, T, Xi, Xt, SYS[sys_name])
    assert_parity(F, G, Fo, Go)
    assert_parity(F0, None, Fo, None)


@pytest.mark.parametrize("scale,T", [(4.0, 2.0), (20.0, 3.0)])
@pytest.mark.parametrize("gradient", ["first_order", "exact"])
def test_large_norm_needs_squarings(scale, T, gradient):
    """||dt*H|| well above the Taylor threshold: scaling-and-squaring (and its derivative) is exercised."""
    D, K, N = 8, 2, 5
    A, B, Xi, Xt = random_system(D, K, seed=7, scale=scale, unitary_targets=True)
    x = np.random.default_rng(3).uniform(-1, 1, (K, N))
    F, G, _ = _run([(A, B, Xi, Xt)], None, T, N, orc.UNITARY_GATE, x, gradient)
    if gradient == "exact":
        Fo, Go = orc.exact_fom_and_gradient(A, B, x, T, Xi, Xt, orc.UNITARY_GATE)
    else:
        Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, orc.UNITARY_GATE)
    assert_parity(F, G, Fo, Go, ftol=1e-9, gtol=1e-7)   # condition grows with ||G||; still far below 1e-6


@pytest.mark.parametrize("D,M", [(2, 5), (2, 8), (4, 3), (4, 6), (8, 5), (16, 2)])
@pytest.mark.parametrize("sys_name", ["state", "unitary"])
@pytest.mark.parametrize("gradient", ["first_order", "exact"])
def test_ensemble(D, M, sys_name, gradient):
    """Weighted ensemble reduction (solve.jl:164-196), including member counts that do not fill a warp."""
    K, N, T = 2, 6, 1.1
    members = [random_system(D, K, seed=300 + 10 * D + k, unitary_targets=(sys_name == "unitary")) for k in range(M)]
    wts = np.random.default_rng(M).random(M)
    x = np.random.default_rng(D * M).uniform(-1, 1, (K, N))
    F, G, F0 = _run(members, wts, T, N, SYS[sys_name], x, gradient)
    if gradient == "exact":
        Fo, Go = orc.ensemble_exact(members, wts, x, T, SYS[sys_name])
    else:
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, x, T, SYS[sys_name])
    assert_parity(F, G, Fo, Go)
    assert_parity(F0, None, Fo, None)


@pytest.mark.parametrize("D,M,R", [(2, 1, 7), (2, 1, 8), (4, 1, 3), (8, 1, 4), (4, 2, 3), (2, 3, 2), (16, 1, 2)])
def test_multistart_batch(D, M, R):
    """R independent pulses per call (multi-start): each must equal the single-pulse result."""
    K, N, T = 2, 8, 1.0
    members = [random_system(D, K, seed=400 + D + k) for k in range(M)]
    wts = np.full(M, 1.0 / M)
    xs = np.random.default_rng(R).uniform(-1, 1, (R, K, N))
    with qoc.GrapeEvaluator(members, T, N, orc.STATE_TRANSFER, wts=wts, n_pulses=R) as ev:
        F, G = ev.eval(xs)
    for r in range(R):
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, xs[r], T, orc.STATE_TRANSFER)
        assert_parity(F[r], G[r], Fo, Go)


@pytest.mark.parametrize("D", [2, 4, 8, 16])
def test_propagators_and_total(D):
    """pw_prop_save! / pw_ham_save! / pw_gen_save! / pw_evolve parity."""
    K, N, T = 3, 6, 1.7
    A, B, _, _ = random_system(D, K, seed=500 + D)
    x = np.random.default_rng(9).uniform(-1, 1, (K, N))
    dt = T / N
    P = qoc.pw_prop_save(A, B, x, K, N, dt)
    Po = orc.pw_prop_save(A, B, x, dt)
    for a, b in zip(P, Po):
        assert np.max(np.abs(a - b)) < 1e-13
    H = qoc.pw_ham_save(A, B, x, K, N)
    for a, b in zip(H, orc.pw_ham_save(A, B, x)):
        assert np.max(np.abs(a - b)) < 1e-14
    Gs = qoc.pw_gen_save(A, B, x, K, N, T)
    for a, b in zip(Gs, orc.pw_gen_save(A, B, x, T)):
        assert np.max(np.abs(a - b)) < 1e-14
    U0 = np.linalg.qr(np.random.default_rng(1).standard_normal((D, D)))[0].astype(np.complex128)
    U = qoc.pw_evolve(A, B, x, K, dt, N, U0)
    assert np.max(np.abs(U - orc.pw_evolve(A, B, x, dt, U0))) < 1e-12


def test_pi_pulse_known_answer():
    """Resonant pi pulse: x = (pi/T, 0), A = 0, B = [Sx, Sy]: rho -> |1><1| exactly, fom 0.75, gradient 0;
    as UnitaryGate with Xt = sigma_x: tau = 2i, fom -4, gradient 0."""
    Sx = np.array([[0, 1], [1, 0]], dtype=complex) / 2
    Sy = np.array([[0, -1j], [1j, 0]], dtype=complex) / 2
    Z = np.zeros((2, 2), dtype=complex)
    N, T = 10, 1.0
    x = np.zeros((2, N)); x[0] = np.pi / T
    r0 = np.diag([1, 0]).astype(complex); r1 = np.diag([0, 1]).astype(complex)
    with qoc.GrapeEvaluator([(Z, [Sx, Sy], r0, r1)], T, N, orc.STATE_TRANSFER) as ev:
        F, G = ev.eval(x)
    assert abs(F - 0.75) < 1e-13 and np.max(np.abs(G)) < 1e-13
    with qoc.GrapeEvaluator([(Z, [Sx, Sy], np.eye(2, dtype=complex), 2 * Sx)], T, N, orc.UNITARY_GATE) as ev:
        F, G = ev.eval(x)
    assert abs(F + 4.0) < 1e-12 and np.max(np.abs(G)) < 1e-12


def test_identical_members_equal_one():
    """Two identical members with weights (1/2, 1/2) equal one member."""
    A, B, Xi, Xt = random_system(4, 2, seed=11)
    x = np.random.default_rng(2).uniform(-1, 1, (2, 12))
    F1, G1, _ = _run([(A, B, Xi, Xt)], None, 1.0, 12, orc.STATE_TRANSFER, x, "first_order")
    F2, G2, _ = _run([(A, B, Xi, Xt)] * 2, [0.5, 0.5], 1.0, 12, orc.STATE_TRANSFER, x, "first_order")
    assert abs(F1 - F2) < 1e-14 and np.max(np.abs(G1 - G2)) < 1e-14


@pytest.mark.parametrize("cfg_name,kw", [("config1", {}), ("config1", {"variant": "readme"}),
                                         ("config2", {"N": 50}), ("config3", {"N": 40}),
                                         ("config4", {"N": 25, "grid": 3})])
def test_baseline_configs_reduced(cfg_name, kw):
    """The BASELINE.json workloads at sizes the oracle finishes in seconds."""
    cfg = getattr(qoc.configs, cfg_name)(**kw)
    F, G, _ = _run(cfg["members"], cfg["wts"], cfg["T"], cfg["N"], cfg["sys_type"], cfg["x"], cfg["gradient"])
    wts = cfg["wts"] if cfg["wts"] is not None else [1.0]
    if cfg["gradient"] == "exact":
        Fo, Go = orc.ensemble_exact(cfg["members"], wts, cfg["x"], cfg["T"], cfg["sys_type"])
    else:
        Fo, Go = orc.ensemble_fom_and_gradient(cfg["members"], wts, cfg["x"], cfg["T"], cfg["sys_type"])
    assert_parity(F, G, Fo, Go)


def test_errors_are_loud():
    A, B, Xi, Xt = random_system(4, 2, seed=1)
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], 1.0, 5, orc.STATE_TRANSFER) as ev:
        with pytest.raises(ValueError):
            ev.eval(np.zeros((3, 5)))
    with pytest.raises(qoc.QocError):
        qoc.GrapeEvaluator([(A, B, Xi, Xt)], 1.0, 0, orc.STATE_TRANSFER)


def test_huge_multistart_batch():
    """R > 65535 pulses in one call (grid-dimension limits) — spot-check a few against the oracle."""
    D, K, N, T, R = 2, 2, 3, 1.0, 70001
    A, B, Xi, Xt = random_system(D, K, seed=77)
    xs = np.random.default_rng(0).uniform(-1, 1, (R, K, N))
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], T, N, orc.STATE_TRANSFER, n_pulses=R) as ev:
        F, G = ev.eval(xs)
    for r in (0, 1, 2, 3, 65535, 65536, R - 2, R - 1):
        Fo, Go = orc.fom_and_gradient_grape(A, B, xs[r], T, Xi, Xt, orc.STATE_TRANSFER)
        assert_parity(F[r], G[r], Fo, Go)


@pytest.mark.parametrize("env", [{"QOC_PHASED": "1", "QOC_CHUNKS": "1"}, {"QOC_PHASED": "1", "QOC_CHUNKS": "4"},
                                 {"QOC_PHASED": "0", "QOC_CHUNKED": "0", "QOC_HAVE_P": "1"},
                                 {"QOC_PHASED": "0", "QOC_CHUNKED": "0", "QOC_HAVE_P": "0"},
                                 {"QOC_CHUNKED": "1", "QOC_CHUNKS": "2"}, {"QOC_CHUNKED": "1", "QOC_CHUNKS": "5"}])
@pytest.mark.parametrize("D,sys_name,gradient", [(8, "unitary", "first_order"), (4, "state", "first_order"),
                                                 (16, "coherence", "first_order"), (4, "unitary", "exact"),
                                                 (8, "state", "exact"), (2, "state", "first_order")])
def test_pipeline_modes(monkeypatch, env, D, sys_name, gradient):
    """Every execution strategy (fused warp-per-chain, slice-parallel exponentials, chunked prefix scan over
    slices with 1 or several chunks) must give the same numbers."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    K, N, T, M = 3, 13, 1.1, 3
    members = [random_system(D, K, seed=900 + D + k, hermitian=(sys_name != "coherence"),
                             unitary_targets=(sys_name == "unitary")) for k in range(M)]
    wts = [0.2, 0.5, 0.3]
    x = np.random.default_rng(D).uniform(-1, 1, (K, N))
    F, G, F0 = _run(members, wts, T, N, SYS[sys_name], x, gradient)
    if gradient == "exact":
        Fo, Go = orc.ensemble_exact(members, wts, x, T, SYS[sys_name])
    else:
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, x, T, SYS[sys_name])
    assert_parity(F, G, Fo, Go)
    assert_parity(F0, None, Fo, None)


@pytest.mark.parametrize("D,K,N,M,R", [(1, 1, 1, 1, 1), (2, 1, 1, 1, 1), (8, 1, 1, 2, 1), (3, 2, 2, 1, 3), (16, 1, 1, 1, 1),
                                       (5, 9, 4, 1, 1), (8, 17, 3, 2, 1), (4, 3, 64, 1, 1), (8, 2, 70, 160, 1)])
@pytest.mark.parametrize("sys_name", ["state", "unitary"])
def test_edge_shapes(D, K, N, M, R, sys_name):
    """Degenerate and ragged shapes: single slice / control / dimension 1, K > 8 (gradient reduction tail), chain
    counts on both sides of the strategy thresholds."""
    T = 0.7
    members = [random_system(D, K, seed=40 + k + D, unitary_targets=(sys_name == "unitary")) for k in range(M)]
    wts = np.linspace(0.5, 1.5, M)
    xs = np.random.default_rng(N).uniform(-1, 1, (R, K, N))
    with qoc.GrapeEvaluator(members, T, N, SYS[sys_name], wts=wts, n_pulses=R) as ev:
        F, G = ev.eval(xs if R > 1 else xs[0])
    for r in range(R):
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, xs[r], T, SYS[sys_name])
        assert_parity(F[r] if R > 1 else F, G[r] if R > 1 else G, Fo, Go)


@pytest.mark.parametrize("D", [4, 8, 64])
def test_no_controls(D):
    """K = 0: drift-only evolution, figure of merit only (empty gradient)."""
    A, _, Xi, Xt = random_system(D, 1, seed=3)
    N, T = 5, 0.9
    x = np.zeros((0, N))
    with qoc.GrapeEvaluator([(A, [], Xi, Xt)], T, N, orc.STATE_TRANSFER) as ev:
        F, G = ev.eval(x)
    Fo, Go = orc.fom_and_gradient_grape(A, [], x, T, Xi, Xt, orc.STATE_TRANSFER)
    assert abs(F - Fo) < 1e-12 and G.shape == (0, N)


def test_c_abi_error_paths_and_stats():
    """Status codes and messages through the raw C ABI (no Python-side validation in between)."""
    import ctypes as C
    lib = qoc._lib.load()
    A, B, Xi, Xt = random_system(4, 2, seed=5)
    h = C.c_void_p()
    d = qoc._lib.QocDesc(sys_type=0, D=4, K=2, N=6, M=1, R=1, T=1.0, gradient=0, convention=0, device=0, expm_theta=0.0, flags=0)
    assert lib.qoc_create(C.byref(h), C.byref(d)) == 0
    x = np.zeros(12); F = np.zeros(1); G = np.zeros(12)
    assert lib.qoc_eval(h, x.ctypes.data, F.ctypes.data, G.ctypes.data) == qoc._lib.QOC_EINVAL      # no system yet
    assert b"qoc_set_system" in lib.qoc_last_error(h)
    cm = lambda M_: np.ascontiguousarray(np.swapaxes(np.asarray(M_, dtype=complex), -1, -2))
    a, b, xi, xt = cm(A), cm(B), cm(Xi), cm(Xt)
    assert lib.qoc_set_system(h, None, b.ctypes.data, xi.ctypes.data, xt.ctypes.data, None, 0) == qoc._lib.QOC_EINVAL
    assert lib.qoc_set_system(h, a.ctypes.data, b.ctypes.data, xi.ctypes.data, xt.ctypes.data, None, 0) == 0
    assert lib.qoc_eval(h, None, F.ctypes.data, G.ctypes.data) == qoc._lib.QOC_EINVAL
    assert lib.qoc_eval(h, x.ctypes.data, F.ctypes.data, None) == 0                                  # value only
    assert lib.qoc_eval(h, x.ctypes.data, F.ctypes.data, G.ctypes.data) == 0
    st = qoc._lib.QocStats()
    assert lib.qoc_get_stats(h, C.byref(st)) == 0
    assert st.n_evals == 2 and st.launches_last_eval >= 2 and st.path == 1 and st.workspace_bytes > 0   # chain kernel + reduction
    bad = qoc._lib.QocDesc(sys_type=0, D=4, K=2, N=6, M=1, R=1, T=1.0, device=99)
    h2 = C.c_void_p()
    assert lib.qoc_create(C.byref(h2), C.byref(bad)) == qoc._lib.QOC_EINVAL and not h2.value
    assert lib.qoc_eval_allreduce_device(h, 1, 1, 1, None) == qoc._lib.QOC_EINVAL                   # comm not connected
    assert lib.qoc_destroy(h) == 0
