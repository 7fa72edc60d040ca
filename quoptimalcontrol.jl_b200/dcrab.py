"""Gradient-free caller of the path: dCRAB (/root/reference/src/dCRAB.jl:13-89) on top of the batched fidelity-only
evaluation (qoc_eval with G = NULL and R > 1).

The reference runs one Nelder-Mead search per super-iteration and evaluates `user_func` once per candidate pulse, serially.
Here every Nelder-Mead step hands ALL the candidates it may need (reflection, expansion, both contractions; the whole
simplex at start-up and on a shrink) to the GPU in one call, so the number of host round trips per step is one instead of
one to four.  The search itself (coefficients alpha = 1, gamma = 2, rho = 1/2, sigma = 1/2, the acceptance rules) is the
textbook method Optim.NelderMead implements; only the order in which function values become available changes, not which
points are accepted."""
from __future__ import annotations

import numpy as np


def nelder_mead_batched(fbatch, x0, initial_step=0.5, max_iters=200, f_tol=1e-10, x_tol=1e-10):
    """Minimise f over R^n.  fbatch(X[m, n]) -> f[m] evaluates m candidates in one call (m <= n + 1).
    Returns (x_best, f_best, iterations, f_calls, batches)."""
    x0 = np.asarray(x0, dtype=np.float64)
    n = x0.size
    simplex = np.vstack([x0] + [x0 + initial_step * np.eye(n)[i] for i in range(n)])
    fvals = np.asarray(fbatch(simplex), dtype=np.float64)
    calls, batches, it = n + 1, 1, 0
    for it in range(1, max_iters + 1):
        order = np.argsort(fvals, kind="stable")
        simplex, fvals = simplex[order], fvals[order]
        if abs(fvals[-1] - fvals[0]) <= f_tol and np.max(np.abs(simplex[1:] - simplex[0])) <= x_tol:
            break
        centroid = simplex[:-1].mean(axis=0)
        worst = simplex[-1]
        xr = centroid + (centroid - worst)             # reflection
        xe = centroid + 2.0 * (centroid - worst)       # expansion
        xoc = centroid + 0.5 * (centroid - worst)      # outside contraction
        xic = centroid - 0.5 * (centroid - worst)      # inside contraction
        fr, fe, foc, fic = fbatch(np.vstack([xr, xe, xoc, xic]))
        calls += 4
        batches += 1
        if fvals[0] <= fr < fvals[-2]:
            simplex[-1], fvals[-1] = xr, fr
        elif fr < fvals[0]:
            simplex[-1], fvals[-1] = (xe, fe) if fe < fr else (xr, fr)
        elif fr < fvals[-1] and foc <= fr:
            simplex[-1], fvals[-1] = xoc, foc
        elif fr >= fvals[-1] and fic < fvals[-1]:
            simplex[-1], fvals[-1] = xic, fic
        else:                                          # shrink towards the best vertex: n new points, one batch
            simplex[1:] = simplex[0] + 0.5 * (simplex[1:] - simplex[0])
            fvals[1:] = fbatch(simplex[1:])
            calls += n
            batches += 1
    best = int(np.argmin(fvals))
    return simplex[best], float(fvals[best]), it, calls, batches


class BatchedFidelity:
    """Adapts a GrapeEvaluator created with n_pulses = R to `fbatch`: pads the candidate list to R pulses per call."""

    def __init__(self, evaluator):
        self.ev, self.R = evaluator, evaluator.R
        self.calls = 0

    def __call__(self, pulses):
        pulses = np.asarray(pulses, dtype=np.float64)
        out = []
        for i in range(0, len(pulses), self.R):
            chunk = pulses[i:i + self.R]
            pad = np.concatenate([chunk, np.repeat(chunk[-1:], self.R - len(chunk), axis=0)]) if len(chunk) < self.R else chunk
            out.append(self.ev.eval_values(pad)[:len(chunk)])
            self.calls += 1
        return np.concatenate(out)


def dCRAB(n_pulses, dt, timeslices, duration, n_freq, n_coeff, initial_guess, user_func=None, batched=None, rng=None,
          max_iters=200):
    """dCRAB as in src/dCRAB.jl:13-89: per super-iteration i a random frequency per pulse, a Fourier ansatz
    coeffs[1] cos(w t) + coeffs[2] sin(w t) added to every pulse (n_coeff = 2 coefficients per pulse), Nelder-Mead over
    the n_coeff * n_pulses coefficients, pulses updated with the minimiser.

    user_func(pulses[K, N]) -> infidelity (serial, like the reference), or batched(pulses[m, K, N]) -> f[m] (e.g.
    BatchedFidelity around a GrapeEvaluator).  Returns (optimised_coeffs, pulses, optim_results)."""
    if n_coeff != 2:
        raise ValueError("the reference's ansatz has two coefficients per pulse (cos, sin)")
    rng = np.random.default_rng() if rng is None else rng
    pulses = np.array(initial_guess, dtype=np.float64).reshape(n_pulses, timeslices)
    t = np.arange(timeslices) * dt                      # pulse_time = 0:dt:duration-dt
    init_freq = rng.random((n_freq, n_pulses))
    init_coeffs = rng.random((n_freq, n_coeff, n_pulses))

    def ansatz(coeffs, freqs):
        c = np.asarray(coeffs).reshape(n_pulses, n_coeff)
        return c[:, :1] * np.cos(freqs[:, None] * t[None]) + c[:, 1:] * np.sin(freqs[:, None] * t[None])

    if batched is None:
        if user_func is None:
            raise TypeError("dCRAB needs user_func or batched")
        batched = lambda P: np.array([user_func(p) for p in P])     # noqa: E731
    optimised_coeffs, optim_results = [], []
    for i in range(n_freq):
        freqs = init_freq[i]
        fbatch = lambda X: batched(np.stack([pulses + ansatz(x, freqs) for x in X]))     # noqa: E731
        x, f, iters, calls, batches = nelder_mead_batched(fbatch, init_coeffs[i].T.reshape(-1), max_iters=max_iters)
        pulses = pulses + ansatz(x, freqs)
        optimised_coeffs.append(x)
        optim_results.append({"minimum": f, "minimizer": x, "iterations": iters, "f_calls": calls, "batches": batches})
    return optimised_coeffs, pulses, optim_results
