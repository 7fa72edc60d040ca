"""world_size-2 gloo test (CPU) of the ensemble sharding + all-reduce plumbing used on N > 1 GPUs.
The local evaluator is a CPU stand-in here (the oracle); on the GPU box it is GrapeEvaluator."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _OracleLocal:
    def __init__(self, members, wts, T, sys_type):
        self.m, self.w, self.T, self.s = members, wts, T, sys_type

    def eval(self, x):
        from oracle import grape_oracle as orc
        return orc.ensemble_fom_and_gradient(self.m, self.w, x, self.T, self.s)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import quoptimalcontrol_jl_b200 as qoc
    from oracle import grape_oracle as orc
    from conftest import random_system
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    M, K, N, T = 5, 2, 6, 1.0
    members = [random_system(4, K, seed=70 + k) for k in range(M)]
    wts = np.linspace(0.1, 0.3, M)
    x = np.random.default_rng(0).uniform(-1, 1, (K, N))
    ev = qoc.ShardedEnsembleEvaluator(members, wts, lambda m, w: _OracleLocal(m, w, T, orc.STATE_TRANSFER), dist=dist)
    F, G = ev.eval(x)
    Fo, Go = orc.ensemble_fom_and_gradient(members, wts, x, T, orc.STATE_TRANSFER)
    q.put((rank, ev.lo, ev.hi, abs(F - Fo), float(np.max(np.abs(G - Go)))))
    dist.destroy_process_group()


class _OracleSliceLocal:
    """CPU stand-in for a GrapeEvaluator that owns a range of slices (total_propagator / set_states / eval_continue)."""
    def __init__(self, A, B, n_slices, duration, sys_type):
        self.A, self.B, self.N, self.T, self.s = A, B, n_slices, duration, sys_type
        self.x = self.Xi = self.Xt = None

    def total_propagator(self, x):
        from oracle import grape_oracle as orc
        self.x = np.asarray(x)
        return orc.pw_evolve(self.A, self.B, self.x, self.T / self.N, np.eye(self.A.shape[0], dtype=complex))

    def set_states(self, Xi, Xt):
        self.Xi, self.Xt = Xi, Xt

    def eval_continue(self):
        from oracle import grape_oracle as orc
        return orc.fom_and_gradient_grape(self.A, self.B, self.x, self.T, self.Xi, self.Xt, self.s)


def _slice_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import quoptimalcontrol_jl_b200 as qoc
    from oracle import grape_oracle as orc
    from conftest import random_system
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    out = []
    for sys_type, unitary_targets in [(orc.STATE_TRANSFER, False), (orc.UNITARY_GATE, True), (orc.COHERENCE_TRANSFER, False)]:
        K, N, T = 2, 7, 1.3                                  # 7 slices over 3 ranks: ranges of 2, 2, 3
        A, B, Xi, Xt = random_system(6, K, seed=90 + sys_type, hermitian=(sys_type != orc.COHERENCE_TRANSFER),
                                     unitary_targets=unitary_targets)
        x = np.random.default_rng(sys_type).uniform(-1, 1, (K, N))
        ev = qoc.SliceParallelEvaluator(Xi, Xt, T, N, sys_type == orc.UNITARY_GATE,
                                        lambda n, dur: _OracleSliceLocal(A, B, n, dur, sys_type), dist=dist)
        F, G = ev.eval(x)
        Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, sys_type)
        out.append((abs(F - Fo) / max(1.0, abs(Fo)), float(np.max(np.abs(G - Go))), float(np.max(np.abs(Go)))))
    q.put((rank, out))
    dist.destroy_process_group()


def test_slice_parallel_world3():
    """Time slices sharded over 3 ranks: range propagators are exchanged, every rank finishes its own slices; the result
    equals the single-process evaluation for all three problem types (both static-sign conventions share the algebra)."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_slice_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, out in res:
        for dF, dG, gmax in out:
            assert dF < 1e-12 and dG < 1e-12 * max(1.0, gmax)


def test_sharded_ensemble_allreduce_world2():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (res[0][1], res[0][2], res[1][1], res[1][2]) == (0, 2, 2, 5)
    for r in res:
        assert r[3] < 1e-14 and r[4] < 1e-14
