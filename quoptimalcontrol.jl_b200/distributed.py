"""Multi-GPU sharding: one process per GPU.

ShardedEnsembleEvaluator: ensemble members block-partitioned over ranks, one all-reduce(sum) of the weighted partial
[F | G] per evaluation — the only collective.  The serial member loop of the reference
(/root/reference/src/solve.jl:166) is the axis that is sharded.

SliceParallelEvaluator: ONE large instance (M = 1), the time slices block-partitioned over ranks; the only exchange is
an all-gather of one D x D range propagator per rank, plus the gather of the gradient blocks."""
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



class SliceParallelEvaluator:
    """Slice-parallel evaluation of one problem (A, B, Xi, Xt) over the ranks of `dist`.

    Rank r owns the slices [lo_r, hi_r) and a local evaluator for that range (N_r slices, duration T N_r / N).  Per call:
      1. U_r = local.total_propagator(x[:, lo:hi])                 (propagators of the range stay on the device)
      2. all-gather of the U_r                                      (the exchange: one D x D matrix per rank)
      3. boundary operators of the range, with L = U_{r-1} ... U_0 and R = U_{n-1} ... U_{r+1}:
           UnitaryGate:  S_lo = L Xi,     C_hi = R' Xt          (src/GRAPE.jl:226, :228 applied to whole ranges)
           density:      S_lo = L Xi L',  C_hi = R' Xt R        (src/GRAPE.jl:245-249)
      4. local.set_states(S_lo, C_hi); F, G_r = local.eval_continue()
      5. all-gather of the gradient blocks.
    Every rank's F is the full figure of merit (tr(S_t' C_t) does not depend on t) and G_r holds exactly the entries the
    single-device evaluation produces for those slices, so the result equals GrapeEvaluator.eval on one device.

    `make_local(n_slices, duration)` builds the local evaluator: a GrapeEvaluator with pure_state=False on a GPU; tests
    inject a CPU stand-in with the same four methods to exercise the plumbing over gloo."""

    def __init__(self, Xi, Xt, T, n_slices, unitary, make_local, dist=None, device=None):
        self.dist = dist
        self.device = device          # torch device for the small boundary products (None: numpy on the host)
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        if n_slices < self.world:
            raise ValueError("SliceParallelEvaluator needs at least one slice per rank")
        self.N, self.unitary = int(n_slices), bool(unitary)
        self.Xi, self.Xt = np.asarray(Xi, dtype=np.complex128), np.asarray(Xt, dtype=np.complex128)
        self.bounds = [shard_bounds(self.N, r, self.world) for r in range(self.world)]
        self.lo, self.hi = self.bounds[self.rank]
        self.local = make_local(self.hi - self.lo, float(T) * (self.hi - self.lo) / self.N)

    def _all_gather(self, arr):
        if self.dist is None or self.world == 1:
            return [arr]
        import torch
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"       # NCCL moves device tensors only
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.float64).reshape(-1).copy()).to(dev)
        sizes = None
        if arr.ndim == 2 and arr.dtype == np.float64:        # gradient blocks: ranks own different numbers of slices
            sizes = [(hi - lo) * arr.shape[0] for lo, hi in self.bounds]
        if sizes is None:
            out = [torch.empty_like(t) for _ in range(self.world)]
            self.dist.all_gather(out, t)
            return [o.cpu().numpy().view(arr.dtype).reshape(arr.shape) for o in out]
        pad = max(sizes)
        tp = torch.zeros(pad, dtype=torch.float64, device=dev)
        tp[:t.numel()] = t
        out = [torch.empty_like(tp) for _ in range(self.world)]
        self.dist.all_gather(out, tp)
        return [o.cpu().numpy()[:n].reshape(arr.shape[0], -1) for o, n in zip(out, sizes)]

    def _boundary_operators(self, U):
        """State before and costate after this rank's range from the gathered range propagators (a handful of D x D
        products: plumbing around the path, done with torch on the device when one is given, else numpy)."""
        if self.device is not None:
            import torch
            to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
            U, Xi, Xt = [to(u) for u in U], to(self.Xi), to(self.Xt)
            mm, dag, back = torch.matmul, (lambda m: m.conj().transpose(-1, -2)), (lambda m: m.cpu().numpy())
        else:
            U, Xi, Xt = list(U), self.Xi, self.Xt
            mm, dag, back = np.matmul, (lambda m: m.conj().T), (lambda m: m)
        L = Rm = None
        for r in range(self.rank):
            L = U[r] if L is None else mm(U[r], L)
        for r in range(self.rank + 1, self.world):
            Rm = U[r] if Rm is None else mm(U[r], Rm)
        S_lo = Xi if L is None else (mm(L, Xi) if self.unitary else mm(mm(L, Xi), dag(L)))
        C_hi = Xt if Rm is None else (mm(dag(Rm), Xt) if self.unitary else mm(mm(dag(Rm), Xt), Rm))
        return back(S_lo), back(C_hi)

    def eval(self, x):
        x = np.asarray(x, dtype=np.float64)
        K = x.shape[0]
        U = self._all_gather(self.local.total_propagator(x[:, self.lo:self.hi]))
        S_lo, C_hi = self._boundary_operators(U)
        self.local.set_states(S_lo, C_hi)
        F, G_r = self.local.eval_continue()
        blocks = self._all_gather(np.ascontiguousarray(G_r, dtype=np.float64).reshape(K, self.hi - self.lo))
        return F, np.concatenate(blocks, axis=1)


class NativeSliceParallelEvaluator:
    """Slice-parallel evaluation of one problem with everything after the pulse upload inside the library
    (qoc_slice_export / qoc_slice_connect / qoc_eval_slice): one process per GPU, the range propagators are exchanged
    through CUDA-IPC peer buffers and the boundary operators are formed by the library's own GEMM kernel reading the peers'
    propagators over NVLink.  `dist` (any torch.distributed backend) only carries the 64-byte IPC handles once.

    eval(x) takes the FULL pulse x[K, N] (every rank passes the same) and returns (F, G_local[K, N_r]); gather() assembles
    the full gradient on every rank (for callers that want it everywhere; the optimiser only needs it on one)."""

    def __init__(self, A, B, Xi, Xt, T, n_slices, sys_type, dist=None, device=0, gradient="first_order", convention="inplace"):
        from .evaluator import GrapeEvaluator
        self.dist = dist
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.N = int(n_slices)
        if self.N < self.world:
            raise ValueError("NativeSliceParallelEvaluator needs at least one slice per rank")
        self.bounds = [shard_bounds(self.N, r, self.world) for r in range(self.world)]
        self.lo, self.hi = self.bounds[self.rank]
        D = np.asarray(A).shape[0]
        I = np.eye(D, dtype=np.complex128)
        self.local = GrapeEvaluator([(A, B, I, I)], float(T) * (self.hi - self.lo) / self.N, self.hi - self.lo, sys_type,
                                    gradient=gradient, convention=convention, device=device, pure_state=False)
        mine = self.local.slice_export()
        handles = [mine]
        if dist is not None and self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, mine)
        self.local.slice_connect(self.world, self.rank, handles, Xi, Xt)

    def eval(self, x, want_grad=True):
        x = np.asarray(x, dtype=np.float64)
        return self.local.eval_slice(np.ascontiguousarray(x[:, self.lo:self.hi]), want_grad)

    def gather(self, G_local):
        if self.dist is None or self.world == 1:
            return G_local
        blocks = [None] * self.world
        self.dist.all_gather_object(blocks, G_local)
        return np.concatenate(blocks, axis=1)

    def close(self):
        self.local.close()
