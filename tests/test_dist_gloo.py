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
