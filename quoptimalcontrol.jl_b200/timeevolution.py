"""GPU versions of the reference's time-evolution helpers (/root/reference/src/timeevolution.jl), same names and
argument order.  Each call builds a throw-away handle; use GrapeEvaluator for repeated calls."""
from __future__ import annotations

import numpy as np

from . import _lib
from .evaluator import GrapeEvaluator


def _ev(A, B, n_slices, T, device):
    D = np.asarray(A).shape[0]
    I = np.eye(D, dtype=np.complex128)
    return GrapeEvaluator([(A, B, I, I)], T, n_slices, _lib.UNITARY_GATE, device=device)


def pw_evolve(A, B, x, n_pulses, dt, n_slices, U0, device=0):
    """U = P_N ... P_1 U0 (src/timeevolution.jl:28-39)."""
    with _ev(A, B, n_slices, dt * n_slices, device) as ev:
        return ev.total_propagator(x) @ np.asarray(U0, dtype=np.complex128)


def pw_evolve_save(A, B, x, n_pulses, dt, n_slices, device=0):
    """List of slice propagators (src/timeevolution.jl:45-57)."""
    with _ev(A, B, n_slices, dt * n_slices, device) as ev:
        return list(ev.propagators(x, "propagator"))


def pw_prop_save(A, B, x, n_pulses, n_slices, dt, out=None, device=0):
    """pw_prop_save! (src/timeevolution.jl:98-110); fills `out` in place when given."""
    P = pw_evolve_save(A, B, x, n_pulses, dt, n_slices, device=device)
    if out is not None:
        for i in range(n_slices):
            out[i][...] = P[i]
        return out
    return P


def pw_ham_save(A, B, x, n_pulses, n_slices, out=None, device=0):
    """pw_ham_save! (src/timeevolution.jl:64-75)."""
    with _ev(A, B, n_slices, 1.0, device) as ev:
        H = list(ev.propagators(x, "hamiltonian"))
    if out is not None:
        for i in range(n_slices):
            out[i][...] = H[i]
        return out
    return H


def pw_gen_save(A, B, x, n_pulses, n_slices, duration, out=None, device=0):
    """pw_gen_save! (src/timeevolution.jl:80-92)."""
    with _ev(A, B, n_slices, duration, device) as ev:
        Gs = list(ev.propagators(x, "generator"))
    if out is not None:
        for i in range(n_slices):
            out[i][...] = Gs[i]
        return out
    return Gs
