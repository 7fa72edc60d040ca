"""The reference's own 12 end-to-end test sets (/root/reference/test/state_transfer_tests.jl and
unitary_gate_tests.jl, fixtures from test/setup_tests.jl) replayed through solve() on the GPU path, with the
same assertions (`sol.result.minimum - C1(target, target) < tol`).  The guesses are seeded here (the reference uses
unseeded `rand`); the optimiser is SciPy's L-BFGS-B standing in for Optim.LBFGS."""
import numpy as np
import pytest

import quoptimalcontrol_jl_b200 as qoc
from oracle import grape_oracle as orc

pytestmark = pytest.mark.gpu

tol = 1e-6                                                        # setup_tests.jl:2
rho_init = np.array([[1, 0], [0, 0]], dtype=complex)              # :4
rho_fin = np.array([[0, 0], [0, 1]], dtype=complex)               # :5
Sx = np.array([[0, 1], [1, 0]], dtype=complex) / 2                # :10
Sy = np.array([[0, -1j], [1j, 0]], dtype=complex) / 2             # :11
Sz = np.array([[1, 0], [0, -1]], dtype=complex) / 2               # :12
U_init = np.eye(2, dtype=complex)                                 # :21
U_fin = np.array([[0, 1], [1, 0]], dtype=complex)                 # :22


def A_gens(k):                                                    # :31
    return (k - 2.5) / 2.5 * Sz * 5


def B_gens(k):                                                    # :32
    return [Sx, Sy]


def odd_switch(k):                                                # :40-46
    return rho_fin if k % 2 else rho_init


def odd_switch_unitary(k):                                        # :57-63
    return U_fin if k % 2 else U_init


def _guess(K, N, seed):
    return np.random.default_rng(seed).random((K, N))


@pytest.mark.parametrize("alg", [qoc.GRAPE(n_slices=10, isinplace=True), qoc.GRAPE(n_slices=10, isinplace=False),
                                 qoc.ADGRAPE(n_slices=10)], ids=["inplace", "static", "adgrape"])
def test_state_transfer(alg):                                     # state_transfer_tests.jl:4-38, 103-119
    prob = qoc.Problem(B=[Sx, Sy], A=Sz, Xi=rho_init, Xt=rho_fin, T=1.0, n_controls=2, guess=_guess(2, 10, 1),
                       sys_type=qoc.StateTransfer())
    sol = qoc.solve(prob, alg)
    assert sol.result.minimum - orc.C1(rho_fin, rho_fin) < tol
    assert sol.opti_pulses.shape == (2, 10) and sol.problem is prob and sol.alg is alg


@pytest.mark.parametrize("alg,tolx", [(qoc.GRAPE(n_slices=25, isinplace=True), 10), (qoc.GRAPE(n_slices=25, isinplace=False), 10),
                                      (qoc.ADGRAPE(n_slices=25), 1)], ids=["inplace", "static", "adgrape"])
def test_state_transfer_ensemble(alg, tolx):                      # state_transfer_tests.jl:42-100, 124-151
    prob = qoc.Problem(B=[Sx, Sy], A=Sz, Xi=rho_init, Xt=rho_fin, T=5.0, n_controls=2, guess=_guess(2, 25, 2),
                       sys_type=qoc.StateTransfer())
    ens = qoc.EnsembleProblem(prob=prob, n_ens=5, A_g=A_gens, B_g=B_gens, XiG=lambda k: rho_init, XtG=odd_switch,
                              wts=np.ones(5) / 5)
    sol = qoc.solve(ens, alg)
    assert sol.result.minimum - orc.C1(rho_fin, rho_fin) < tol * tolx


@pytest.mark.parametrize("alg", [qoc.GRAPE(n_slices=10, isinplace=True), qoc.GRAPE(n_slices=10, isinplace=False)],
                         ids=["inplace", "static"])
def test_unitary_x_gate(alg):                                     # unitary_gate_tests.jl:3-37
    prob = qoc.Problem(B=[Sx, Sy], A=Sz, Xi=U_init, Xt=U_fin, T=1.0, n_controls=2, guess=_guess(2, 10, 3),
                       sys_type=qoc.UnitaryGate())
    sol = qoc.solve(prob, alg)
    assert sol.result.minimum - orc.C1(U_fin, U_fin) < tol


def test_unitary_x_gate_adgrape():                                # unitary_gate_tests.jl:115-133
    prob = qoc.Problem(B=[Sx, Sy], A=Sz, Xi=U_init, Xt=U_fin, T=1.0, n_controls=2, guess=_guess(2, 25, 4),
                       sys_type=qoc.UnitaryGate())
    sol = qoc.solve(prob, qoc.ADGRAPE(n_slices=25))
    assert sol.result.minimum - orc.C1(U_fin, U_fin) < tol
    # verify the pulse like the reference does (pw_evolve, :130): the gate is X up to a phase
    U = qoc.pw_evolve(Sz, [Sx, Sy], sol.opti_pulses, 2, 1.0 / 25, 25, U_init)
    assert orc.C1(U_fin, U) < 1e-5


@pytest.mark.parametrize("alg,T", [(qoc.GRAPE(n_slices=100, isinplace=True, optim_options={"ftol": 1e-3}), 5.0),
                                   (qoc.GRAPE(n_slices=100, isinplace=False, optim_options={"ftol": 1e-3}), 10.0),
                                   (qoc.ADGRAPE(n_slices=100), 5.0)], ids=["inplace", "static", "adgrape"])
def test_robust_x_gate(alg, T):                                   # unitary_gate_tests.jl:41-112, 137-167
    prob = qoc.Problem(B=[Sx, Sy], A=Sz, Xi=U_init, Xt=U_fin, T=T, n_controls=2, guess=_guess(2, 100, 5),
                       sys_type=qoc.UnitaryGate())
    ens = qoc.EnsembleProblem(prob=prob, n_ens=5, A_g=A_gens, B_g=B_gens, XiG=lambda k: U_init, XtG=odd_switch_unitary,
                              wts=np.ones(5) / 5)
    sol = qoc.solve(ens, alg)
    assert sol.result.minimum - orc.C1(rho_fin, rho_fin) < tol    # the reference compares against C1(rho_fin, rho_fin) here


def test_closure_matches_oracle_during_optimisation():
    """Every (F, G) the optimiser sees equals the oracle's at that pulse (spot check along an optimisation path)."""
    prob = qoc.Problem(B=[Sx, Sy], A=Sz, Xi=rho_init, Xt=rho_fin, T=1.0, n_controls=2, guess=_guess(2, 10, 6),
                       sys_type=qoc.StateTransfer())
    sol = qoc.solve(prob, qoc.GRAPE(n_slices=10, optim_options={"maxiter": 5}))
    x = sol.opti_pulses
    with qoc.GrapeEvaluator([(Sz, [Sx, Sy], rho_init, rho_fin)], 1.0, 10, orc.STATE_TRANSFER) as ev:
        F, G = ev.eval(x)
    Fo, Go = orc.fom_and_gradient_grape(Sz, [Sx, Sy], x, 1.0, rho_init, rho_fin, orc.STATE_TRANSFER)
    assert abs(F - Fo) < 1e-12 and np.max(np.abs(G - Go)) < 1e-12
    assert abs(sol.fidelity - Fo) < 1e-12


@pytest.mark.parametrize("linesearch", ["hagerzhang", "backtracking"])
@pytest.mark.parametrize("kind", ["state", "unitary_adgrape", "ensemble"])
def test_native_lbfgs_solves_reference_scenarios(kind, linesearch):
    """qoc_minimize_lbfgs (the optimiser loop inside the library) on the reference's scenarios, same assertions, with the
    Hager-Zhang line search (Optim.LBFGS's default) and with plain backtracking."""
    import time
    if kind == "state":
        prob = qoc.Problem(B=[Sx, Sy], A=Sz, Xi=rho_init, Xt=rho_fin, T=1.0, n_controls=2, guess=_guess(2, 10, 11), sys_type=qoc.StateTransfer())
        alg = qoc.GPUGRAPE(n_slices=10, optimizer="native", optim_options={"linesearch": linesearch})
        target = orc.C1(rho_fin, rho_fin)
    elif kind == "unitary_adgrape":
        prob = qoc.Problem(B=[Sx, Sy], A=Sz, Xi=U_init, Xt=U_fin, T=1.0, n_controls=2, guess=_guess(2, 25, 12), sys_type=qoc.UnitaryGate())
        alg = qoc.GPUGRAPE(n_slices=25, gradient="exact", optimizer="native", optim_options={"linesearch": linesearch})
        target = orc.C1(U_fin, U_fin)
    else:
        p0 = qoc.Problem(B=[Sx, Sy], A=Sz, Xi=rho_init, Xt=rho_fin, T=5.0, n_controls=2, guess=_guess(2, 25, 13), sys_type=qoc.StateTransfer())
        prob = qoc.EnsembleProblem(prob=p0, n_ens=5, A_g=A_gens, B_g=B_gens, XiG=lambda k: rho_init, XtG=odd_switch, wts=np.ones(5) / 5)
        alg = qoc.GPUGRAPE(n_slices=25, optimizer="native", optim_options={"linesearch": linesearch})
        target = orc.C1(rho_fin, rho_fin)
    t0 = time.perf_counter(); sol = qoc.solve(prob, alg); t_native = time.perf_counter() - t0
    assert sol.result.minimum - target < tol * 10
    assert sol.result.f_calls >= sol.result.iterations and sol.opti_pulses.shape == np.asarray(sol.problem.guess if kind != "ensemble" else p0.guess).shape
    # the SciPy-driven solve reaches the same kind of minimum
    alg2 = qoc.GPUGRAPE(n_slices=alg.n_slices, gradient=alg.gradient, optimizer="scipy")
    t0 = time.perf_counter(); sol2 = qoc.solve(prob, alg2); t_scipy = time.perf_counter() - t0
    assert sol2.result.minimum - target < tol * 10
    print(f"{kind} [{linesearch}]: native {sol.result.iterations} it / {sol.result.f_calls} evals in {t_native*1e3:.1f} ms (min {sol.result.minimum:.6f}); "
          f"scipy {sol2.result.iterations} it / {sol2.result.f_calls} evals in {t_scipy*1e3:.1f} ms (min {sol2.result.minimum:.6f})")
