"""Multi-GPU sharding of the ensemble: one process per GPU, members block-partitioned over
ranks, one all-reduce(sum) of the weighted partial [F | G] per evaluation — the only collective.
The serial member loop of the reference (/root/reference/src/solve.jl:166) is the axis that is sharded."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_members: int, rank: int, world: int):
    """Contiguous block [lo, hi) of members owned by `rank`; blocks differ by at most one member."""
    return rank * n_members // world, (rank + 1) * n_members // world


class ShardedEnsembleEvaluator:
    """Each rank evaluates its members' weighted partial sum; `eval` returns the all-reduced (F, G) on every rank.

    `make_local(members, wts)` builds the local evaluator (GrapeEvaluator on a GPU; tests inject a CPU stand-in
    to exercise the plumbing over gloo).  The local evaluator must expose eval(x) -> (F, G[K, N])."""

    def __init__(self, members, wts, make_local, dist=None):
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        lo, hi = shard_bounds(len(members), self.rank, self.world)
        self.lo, self.hi = lo, hi
        wts = np.ones(len(members)) if wts is None else np.asarray(wts, dtype=np.float64)
        self.local = make_local(members[lo:hi], wts[lo:hi]) if hi > lo else None

    def eval(self, x):
        import torch
        K, N = np.asarray(x).shape
        fg = np.zeros(1 + K * N)
        if self.local is not None:
            F, G = self.local.eval(x)
            fg[0] = F
            fg[1:] = np.asarray(G).ravel()
        if self.dist is not None and self.world > 1:
            t = torch.from_numpy(fg)
            self.dist.all_reduce(t)          # sum over ranks
            fg = t.numpy()
        return float(fg[0]), fg[1:].reshape(K, N)
