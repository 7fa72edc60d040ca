"""The reference's own end-to-end test sets (/root/reference/test/state_transfer_tests.jl, unitary_gate_tests.jl, fixtures
from test/setup_tests.jl) replayed on the CPU through the ORACLE: the only assertions the reference makes about this path
are one-sided bounds on the optimiser's final minimum (`sol.result.minimum - C1(target, target) < tol`), and the oracle's
(F, G) closure meets every one of them.  This is the pinning the reference offers (no stored values exist);
tests/test_gpu_reference_scenarios.py replays the same scenarios through the CUDA path.
SciPy's L-BFGS-B stands in for Optim.LBFGS; the guesses are seeded (the reference uses unseeded `rand`)."""
import numpy as np
import pytest
from scipy.optimize import minimize

from oracle import grape_oracle as orc

tol = 1e-6                                                        # setup_tests.jl:2
rho_init = np.array([[1, 0], [0, 0]], dtype=complex)              # :4
rho_fin = np.array([[0, 0], [0, 1]], dtype=complex)               # :5
Sx = np.array([[0, 1], [1, 0]], dtype=complex) / 2                # :10
Sy = np.array([[0, -1j], [1j, 0]], dtype=complex) / 2             # :11
Sz = np.array([[1, 0], [0, -1]], dtype=complex) / 2               # :12
U_init = np.eye(2, dtype=complex)                                 # :21
U_fin = np.array([[0, 1], [1, 0]], dtype=complex)                 # :22
A_gens = lambda k: (k - 2.5) / 2.5 * Sz * 5                       # :31 (k = 1..n_ens, Julia indexing)
odd_switch = lambda k: rho_fin if k % 2 else rho_init             # :40-46
odd_switch_unitary = lambda k: U_fin if k % 2 else U_init         # :57-63


def _solve(members, wts, K, N, T, sys_type, mode, seed, ftol=None):
    """The (F, G, x) closure of src/solve.jl:75-100 / :164-196 (GRAPE) or :268-361 (ADGRAPE) built from the oracle."""
    def fg(v):
        x = v.reshape(K, N)
        if mode == "adgrape":
            F, G = orc.ensemble_exact(members, wts, x, T, sys_type)
        else:
            F, G = orc.ensemble_fom_and_gradient(members, wts, x, T, sys_type, orc.REF_INPLACE if mode == "inplace" else orc.REF_STATIC)
        return F, np.asarray(G).ravel()
    opts = {"maxiter": 1000, "gtol": 1e-10, "ftol": 1e-15 if ftol is None else ftol}
    return minimize(fg, np.random.default_rng(seed).random(K * N), jac=True, method="L-BFGS-B", options=opts)


def _ensemble(XiG, XtG):
    return [(A_gens(k), [Sx, Sy], XiG(k), XtG(k)) for k in range(1, 6)], np.ones(5) / 5      # init_ensemble, src/tools.jl:42-53


@pytest.mark.parametrize("mode", ["inplace", "static", "adgrape"])
def test_state_transfer(mode):                                    # state_transfer_tests.jl:4-38, 103-119
    res = _solve([(Sz, [Sx, Sy], rho_init, rho_fin)], [1.0], 2, 10, 1.0, orc.STATE_TRANSFER, mode, 1)
    assert res.fun - orc.C1(rho_fin, rho_fin) < tol


@pytest.mark.parametrize("mode,tolx", [("inplace", 10), ("static", 10), ("adgrape", 1)])
def test_state_transfer_ensemble(mode, tolx):                     # state_transfer_tests.jl:42-100, 124-151
    members, wts = _ensemble(lambda k: rho_init, odd_switch)
    res = _solve(members, wts, 2, 25, 5.0, orc.STATE_TRANSFER, mode, 2)
    assert res.fun - orc.C1(rho_fin, rho_fin) < tol * tolx


@pytest.mark.parametrize("mode,N,seed", [("inplace", 10, 3), ("static", 10, 3), ("adgrape", 25, 4)])
def test_unitary_x_gate(mode, N, seed):                           # unitary_gate_tests.jl:3-37, 115-133
    res = _solve([(Sz, [Sx, Sy], U_init, U_fin)], [1.0], 2, N, 1.0, orc.UNITARY_GATE, mode, seed)
    assert res.fun - orc.C1(U_fin, U_fin) < tol
    if mode == "adgrape":                                         # the reference re-evolves the pulse (:130)
        U = orc.pw_evolve(Sz, [Sx, Sy], res.x.reshape(2, N), 1.0 / N, U_init)
        assert orc.C1(U_fin, U) < 1e-5


@pytest.mark.parametrize("mode,T,ftol", [("inplace", 5.0, 1e-3), ("static", 10.0, 1e-3), ("adgrape", 5.0, None)])
def test_robust_x_gate(mode, T, ftol):                            # unitary_gate_tests.jl:41-112, 137-167
    members, wts = _ensemble(lambda k: U_init, odd_switch_unitary)
    res = _solve(members, wts, 2, 100, T, orc.UNITARY_GATE, mode, 5, ftol)
    assert res.fun - orc.C1(rho_fin, rho_fin) < tol               # the reference compares against C1(rho_fin, rho_fin) here
