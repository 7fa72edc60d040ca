"""Algorithm types and solve(), mirroring /root/reference/src/solve.jl.  The fidelity+gradient closure is one
call into libqocgrape.so; the optimiser is SciPy's L-BFGS-B standing in for Optim.LBFGS (Optim.jl is Julia-only;
line search details differ, the closure contract `topt(F, G, x)` is the same)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np
from scipy.optimize import minimize

from .evaluator import GrapeEvaluator
from .problems import EnsembleProblem, Problem, UnitaryGate, init_ensemble


@dataclass
class Piecewise:                       # src/timeevolution.jl:11-14
    n_slices: int = 1
    expm_method: str = "fast"          # stored and ignored, like the reference


@dataclass
class GPUGRAPE:
    """The new algorithm type that plugs into solve(prob, alg) beside GRAPE / ADGRAPE (src/solve.jl:33-52)."""
    n_slices: int
    gradient: str = "first_order"      # "first_order" (GRAPE) | "exact" (ADGRAPE semantics)
    convention: str = "inplace"        # UnitaryGate first-order sign: grad_func! vs grad_func
    device: int = 0
    optim_options: dict = field(default_factory=dict)
    optimizer: str = "scipy"           # "scipy" (L-BFGS-B, host) | "native" (qoc_minimize_lbfgs inside the library)
    pure_state: bool = True            # D > 16 pure-state transfers on sparse closed systems: state-vector sweep (same F, G)
    devices: Any = None                # list of CUDA ordinals: ensemble members sharded over these GPUs inside this process
    penalty: tuple = (0.0, 0.0)        # weights of the C3 / C4 control penalties (src/cost_functions.jl:29-39)

    @property
    def integrator(self):
        return Piecewise(self.n_slices)


def GRAPE(n_slices, expm_method="fast", isinplace=True, optim_options=None, device=0):
    """src/solve.jl:39-42 keyword constructor; `isinplace` selects grad_func! vs grad_func semantics."""
    return GPUGRAPE(n_slices=n_slices, gradient="first_order", convention="inplace" if isinplace else "static",
                    device=device, optim_options=optim_options or {})


def ADGRAPE(n_slices, expm_method="fast", optim_options=None, device=0):
    """src/solve.jl:54-57: exact gradient of the C1 functional."""
    return GPUGRAPE(n_slices=n_slices, gradient="exact", device=device, optim_options=optim_options or {})


@dataclass
class SolutionResult:                  # src/solve.jl:14-20
    result: Any
    fidelity: float
    opti_pulses: np.ndarray
    problem: Any
    alg: Any


EnsembleSolutionResult = SolutionResult    # src/solve.jl:23-29 has the same fields


class _NativeResult:
    """Result of qoc_minimize_lbfgs with the same attribute names."""
    def __init__(self, x, info):
        self.minimum = float(info["minimum"])
        self.minimizer = x
        self.iterations = int(info["iterations"])
        self.f_calls = self.g_calls = int(info["f_calls"])
        self.converged = bool(info["converged"])
        self.g_norm = float(info["g_norm"])
        self.raw = info


class _OptimResult:
    """Subset of Optim.jl's result object used by callers of the reference (minimum, minimizer, counts)."""
    def __init__(self, res, shape):
        self.minimum = float(res.fun)
        self.minimizer = np.asarray(res.x).reshape(shape)
        self.iterations = int(res.nit)
        self.f_calls = int(res.nfev)
        self.g_calls = int(res.nfev)
        self.converged = bool(res.success)
        self.raw = res


def _optimize(ev, guess, options):
    shape = np.asarray(guess).shape

    def topt(xflat):                   # the (F, G, x) closure of src/solve.jl:75-100 / :164-196
        F, G = ev.eval(xflat.reshape(shape), want_grad=True)
        return F, G.ravel()

    opts = {"maxiter": 1000, "gtol": 1e-8, "ftol": 1e-15}
    opts.update({k: v for k, v in (options or {}).items() if k != "linesearch"})
    res = minimize(topt, np.asarray(guess, dtype=np.float64).ravel(), jac=True, method="L-BFGS-B", options=opts)
    return _OptimResult(res, shape)


def solve(prob, alg=None):
    """solve(prob::Problem, alg) (src/solve.jl:63-143, :255-266) and solve(ens::EnsembleProblem, alg)
    (:145-250, :293-310)."""
    if alg is None:
        raise TypeError("solve(prob) needs an algorithm: the reference's default GRAPE() requires n_slices")
    if isinstance(prob, EnsembleProblem):
        members = init_ensemble(prob)
        first = members[0]
        tuples = [(m.A, m.B, m.Xi, m.Xt) for m in members]
        wts, guess = prob.wts, first.guess
    else:
        first = prob
        tuples = [(prob.A, prob.B, prob.Xi, prob.Xt)]
        wts, guess = None, prob.guess
    with GrapeEvaluator(tuples, first.T, alg.n_slices, first.sys_type, wts=wts, gradient=alg.gradient,
                        convention=alg.convention, device=alg.device, pure_state=getattr(alg, "pure_state", True),
                        devices=getattr(alg, "devices", None), penalty=getattr(alg, "penalty", (0.0, 0.0))) as ev:
        if getattr(alg, "optimizer", "scipy") == "native":
            o = alg.optim_options or {}
            x, info = ev.minimize_lbfgs(guess, max_iters=o.get("maxiter", o.get("iterations", 0)), g_tol=o.get("gtol", o.get("g_tol", 0.0)),
                                        f_tol=o.get("ftol", o.get("f_tol", -1.0)), linesearch=o.get("linesearch", "hagerzhang"))
            res = _NativeResult(x, info)
        else:
            res = _optimize(ev, guess, alg.optim_options)
    return SolutionResult(res, res.minimum, res.minimizer, prob, alg)
