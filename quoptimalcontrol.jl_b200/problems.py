"""Problem description types, mirroring /root/reference/src/problems.jl (field names unchanged)."""
from __future__ import annotations

import copy
from dataclasses import dataclass, replace
from typing import Any, Callable, Sequence

import numpy as np

from . import _lib


class SystemType:            # src/problems.jl:1
    code = -1


class ClosedSystem(SystemType):
    pass


class OpenSystem(SystemType):
    pass


class StateTransfer(ClosedSystem):        # src/problems.jl:8
    code = _lib.STATE_TRANSFER


class UnitaryGate(ClosedSystem):          # src/problems.jl:9
    code = _lib.UNITARY_GATE


class CoherenceTransfer(OpenSystem):      # src/problems.jl:10
    code = _lib.COHERENCE_TRANSFER


@dataclass
class Problem:
    """src/problems.jl:19-28."""
    B: Sequence[np.ndarray]      # control terms
    A: np.ndarray                # drift terms
    Xi: np.ndarray               # initial state
    Xt: np.ndarray               # target operator or state
    T: float                     # duration of pulse
    n_controls: int              # number of pulses
    guess: np.ndarray            # guess at controls, shape (n_controls, n_slices)
    sys_type: SystemType


@dataclass
class EnsembleProblem:
    """src/problems.jl:33-41.  The generators take the 1-based member index k like the reference's closures."""
    prob: Problem
    n_ens: int
    A_g: Callable[[int], np.ndarray]
    B_g: Callable[[int], Sequence[np.ndarray]]
    XiG: Callable[[int], np.ndarray]
    XtG: Callable[[int], np.ndarray]
    wts: Sequence[float]


# README.md:51 / examples/examples.jl:12,28,44 — the older problem names, as keyword constructors onto Problem.
def ClosedStateTransfer(**kw) -> Problem:
    return Problem(sys_type=StateTransfer(), **kw)


def UnitarySynthesis(**kw) -> Problem:
    return Problem(sys_type=UnitaryGate(), **kw)


def OpenSystemCoherenceTransfer(**kw) -> Problem:
    return Problem(sys_type=CoherenceTransfer(), **kw)


def init_ensemble(ens: EnsembleProblem):
    """src/tools.jl:42-53: one Problem per member with A, B, Xi, Xt replaced by the generators' values."""
    out = []
    for k in range(1, ens.n_ens + 1):
        p = copy.copy(ens.prob)
        out.append(replace(p, A=ens.A_g(k), B=ens.B_g(k), Xi=ens.XiG(k), Xt=ens.XtG(k)))
    return out
