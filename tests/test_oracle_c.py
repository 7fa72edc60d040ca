"""The C restatement (oracle/grape_oracle.c) against the numpy oracle and scipy's expm."""
import numpy as np
import pytest
from scipy.linalg import expm

from oracle import c_oracle, grape_oracle as orc
from conftest import random_system


@pytest.mark.parametrize("norm,prods", [(0.01, 2), (0.1, 3), (0.5, 4), (1.5, 5), (4.0, 6), (30.0, 9)])
def test_pade_ladder(norm, prods):
    """Julia exp! degree ladder: product counts 2/3/4/5 then 6 + squarings, values equal scipy's expm."""
    rng = np.random.default_rng(1)
    X = rng.standard_normal((6, 6)) + 1j * rng.standard_normal((6, 6))
    X *= norm / np.linalg.norm(X, 1)
    E, p = c_oracle.expm(X)
    assert p == prods
    assert np.max(np.abs(E - expm(X))) < 1e-13 * max(1.0, np.linalg.norm(expm(X), 1))
    assert orc.julia_exp_rung(X)[1] + orc.julia_exp_rung(X)[2] == prods


@pytest.mark.parametrize("sys_type", [orc.STATE_TRANSFER, orc.UNITARY_GATE, orc.COHERENCE_TRANSFER])
@pytest.mark.parametrize("variant", [0, 1])
def test_c_matches_numpy(sys_type, variant):
    D, K, N, M, T = 4, 3, 7, 3, 1.4
    members = [random_system(D, K, seed=30 + k, hermitian=(sys_type != orc.COHERENCE_TRANSFER),
                             unitary_targets=(sys_type == orc.UNITARY_GATE)) for k in range(M)]
    wts = [0.2, 0.3, 0.5]
    x = np.random.default_rng(2).uniform(-1, 1, (K, N))
    Fo, Go = orc.ensemble_fom_and_gradient(members, wts, x, T, sys_type, variant)
    for nthreads in (1, 2):
        F, G = c_oracle.eval_ensemble(members, wts, x, T, sys_type, variant, nthreads)
        assert abs(F - Fo) < 1e-12 * max(1, abs(Fo))
        assert np.max(np.abs(G - Go)) < 1e-12 * max(1, np.max(np.abs(Go)))
