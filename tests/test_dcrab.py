"""Gradient-free surface: batched Nelder-Mead / dCRAB mirror of /root/reference/src/dCRAB.jl on the
CPU (plumbing, serial user_func like the reference) and on the GPU's batched fidelity-only evaluation against the oracle."""
import numpy as np
import pytest

import quoptimalcontrol_jl_b200 as qoc
from oracle import grape_oracle as orc
from conftest import random_system


def test_nelder_mead_batched_minimises_a_quadratic():
    A = np.diag([1.0, 4.0, 0.5])
    c = np.array([0.3, -0.2, 0.7])
    calls = []

    def fbatch(X):
        calls.append(len(X))
        return np.array([(x - c) @ A @ (x - c) for x in X])
    x, f, iters, ncalls, batches = qoc.nelder_mead_batched(fbatch, np.zeros(3), max_iters=400)
    assert np.max(np.abs(x - c)) < 1e-4 and f < 1e-8
    assert batches == len(calls) and ncalls == sum(calls) and max(calls) <= 4


def test_dcrab_serial_user_func_reduces_infidelity():
    """The reference's contract: user_func(pulses[K, N]) -> infidelity, evaluated one candidate at a time."""
    Sx = np.array([[0, 1], [1, 0]], dtype=complex) / 2
    Sy = np.array([[0, -1j], [1j, 0]], dtype=complex) / 2
    Sz = np.diag([0.5, -0.5]).astype(complex)
    rho0, rho1 = np.diag([1, 0]).astype(complex), np.diag([0, 1]).astype(complex)
    N, T = 20, 2.0

    def infid(x):
        return orc.exact_functional(Sz, [Sx, Sy], x, T, rho0, rho1, orc.STATE_TRANSFER)
    guess = np.zeros((2, N))
    f0 = infid(guess)
    coeffs, pulses, res = qoc.dCRAB(2, T / N, N, T, 2, 2, guess, user_func=infid, rng=np.random.default_rng(1), max_iters=150)
    assert pulses.shape == (2, N) and len(coeffs) == 2 and len(res) == 2
    assert res[-1]["minimum"] < f0 - 0.05 and abs(infid(pulses) - res[-1]["minimum"]) < 1e-12


@pytest.mark.gpu
def test_dcrab_on_the_batched_gpu_evaluation():
    """Every Nelder-Mead batch is ONE qoc_eval call (G = NULL, R = 5); the values equal the oracle's functional and the
    search reaches the same kind of minimum as the serial run."""
    D, K, N, T = 4, 2, 24, 3.0
    A, B, _, _ = random_system(D, K, seed=77)
    Xi = np.zeros((D, D), dtype=complex); Xi[0, 0] = 1
    Xt = np.zeros((D, D), dtype=complex); Xt[1, 1] = 1
    guess = np.zeros((K, N))
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], T, N, orc.STATE_TRANSFER, gradient="exact", n_pulses=5) as ev:
        fb = qoc.BatchedFidelity(ev)
        xs = np.random.default_rng(0).uniform(-1, 1, (7, K, N))
        vals = fb(xs)                                   # 7 candidates -> two calls of 5
        for x, v in zip(xs, vals):
            assert abs(v - orc.exact_functional(A, B, x, T, Xi, Xt, orc.STATE_TRANSFER)) < 1e-12
        assert fb.calls == 2
        f0 = float(fb(guess[None])[0])
        coeffs, pulses, res = qoc.dCRAB(K, T / N, N, T, 3, 2, guess, batched=fb, rng=np.random.default_rng(2), max_iters=120)
        assert res[-1]["minimum"] < f0 - 1e-3
        assert abs(orc.exact_functional(A, B, pulses, T, Xi, Xt, orc.STATE_TRANSFER) - res[-1]["minimum"]) < 1e-11
        assert all(r["batches"] <= r["iterations"] * 2 + 1 for r in res)
