"""quoptimalcontrol.jl_b200 — B200-native (sm_100a) GRAPE fidelity+gradient hot path of QuOptimalControl.jl.

Host-side mirror of the reference's Problem / EnsembleProblem / solve(prob, alg) surface; every evaluation runs
in libqocgrape.so (hand-written CUDA, C ABI in include/qocgrape.h).  There is no CPU fallback."""
from . import _lib, configs
from ._lib import QocError
from .build import build
from .distributed import (NativeSliceParallelEvaluator, ShardedEnsembleEvaluator, SliceParallelEvaluator,
                          shard_bounds)
from .dcrab import BatchedFidelity, dCRAB, nelder_mead_batched
from .evaluator import GrapeEvaluator
from .problems import (ClosedStateTransfer, ClosedSystem, CoherenceTransfer, EnsembleProblem,
                       OpenSystem, OpenSystemCoherenceTransfer, Problem, StateTransfer, SystemType,
                       UnitaryGate, UnitarySynthesis, init_ensemble)
from .solve import (ADGRAPE, GPUGRAPE, GRAPE, EnsembleSolutionResult, Piecewise, SolutionResult, solve)
from .timeevolution import pw_evolve, pw_evolve_save, pw_gen_save, pw_ham_save, pw_prop_save

__all__ = [n for n in dir() if not n.startswith("_")]
