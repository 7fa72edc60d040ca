"""Two-GPU test of the sharded ensemble with the fused one-shot NVLink all-reduce (qoc_comm_* / qoc_eval_allreduce_device).
Skipped on single-GPU boxes; run with `gpurun --gpus 2`."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import quoptimalcontrol_jl_b200 as qoc
    from oracle import grape_oracle as orc
    from conftest import random_system
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)   # only carries handles
    M, K, N, T, D = 7, 3, 40, 1.0, 8
    members = [random_system(D, K, seed=80 + k, unitary_targets=True) for k in range(M)]
    wts = np.linspace(0.05, 0.25, M)
    lo, hi = qoc.shard_bounds(M, rank, world)
    ev = qoc.GrapeEvaluator(members[lo:hi], T, N, orc.UNITARY_GATE, wts=wts[lo:hi], device=rank)
    handles = [None] * world
    dist.all_gather_object(handles, ev.comm_export())
    ev.comm_connect(world, rank, handles)
    dev = torch.device("cuda", rank)
    errs = []
    for it in range(5):                                   # several epochs: exercises both halves of the exchange buffer
        x = np.random.default_rng(it).uniform(-1, 1, (K, N))
        x_dev = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
        fg = torch.zeros(N * K + 1, dtype=torch.float64, device=dev)
        ev.eval_allreduce_device(x_dev.data_ptr(), fg.data_ptr(), True, None)
        torch.cuda.synchronize()
        out = fg.cpu().numpy()
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, x, T, orc.UNITARY_GATE)
        G = out[1:].reshape(N, K).T
        errs.append((abs(out[0] - Fo), float(np.max(np.abs(G - Go)) / np.max(np.abs(Go)))))
        gathered = [None] * world
        dist.all_gather_object(gathered, out.tobytes())
        assert all(g == gathered[0] for g in gathered), "ranks disagree bitwise"
    q.put((rank, errs))
    ev.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_oneshot_allreduce_two_gpus():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        for ef, eg in errs:
            assert ef < 1e-10 and eg < 1e-8


def _slice_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import quoptimalcontrol_jl_b200 as qoc
    from oracle import grape_oracle as orc
    from conftest import random_system
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    errs = []
    for sys_type, name in [(orc.STATE_TRANSFER, "state"), (orc.UNITARY_GATE, "unitary"), (orc.COHERENCE_TRANSFER, "coherence")]:
        D, K, N, T = 64, 2, 11, 0.9
        A, B, Xi, Xt = random_system(D, K, seed=120 + sys_type, hermitian=(sys_type != orc.COHERENCE_TRANSFER),
                                     unitary_targets=(sys_type == orc.UNITARY_GATE))
        x = np.random.default_rng(7).uniform(-1, 1, (K, N))
        I = np.eye(D, dtype=complex)
        ev = qoc.SliceParallelEvaluator(
            Xi, Xt, T, N, sys_type == orc.UNITARY_GATE,
            lambda n, dur: qoc.GrapeEvaluator([(A, B, I, I)], dur, n, sys_type, device=rank, pure_state=False), dist=dist,
            device=torch.device("cuda", rank) if sys_type != orc.STATE_TRANSFER else None)     # both host and device boundary products
        for _ in range(2):                       # second call: states replaced again on the same handle
            F, G = ev.eval(x)
        Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, sys_type)
        errs.append((name, abs(F - Fo) / max(1.0, abs(Fo)), float(np.max(np.abs(G - Go)) / max(np.max(np.abs(Go)), 1e-6))))
        ev.local.close()
    q.put((rank, errs))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_slice_parallel_two_gpus():
    """One D = 64 instance, its 11 slices split over two GPUs (qoc_total_propagator -> exchange -> qoc_set_states ->
    qoc_eval_continue): same F and G as the single-device evaluation / the oracle."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_slice_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        for name, ef, eg in errs:
            assert ef < 1e-10 and eg < 1e-8, (rank, name, ef, eg)


def _native_slice_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import quoptimalcontrol_jl_b200 as qoc
    from oracle import grape_oracle as orc
    from conftest import random_system
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)   # carries the IPC handles
    errs = []
    for sys_type, name in [(orc.STATE_TRANSFER, "state"), (orc.UNITARY_GATE, "unitary"), (orc.COHERENCE_TRANSFER, "coherence")]:
        D, K, N, T = 64, 2, 4 * world + 3, 0.9
        A, B, Xi, Xt = random_system(D, K, seed=220 + sys_type, hermitian=(sys_type != orc.COHERENCE_TRANSFER),
                                     unitary_targets=(sys_type == orc.UNITARY_GATE))
        ev = qoc.NativeSliceParallelEvaluator(A, B, Xi, Xt, T, N, sys_type, dist=dist, device=rank)
        for it in range(3):
            x = np.random.default_rng(7 + it).uniform(-1, 1, (K, N))
            F, G_loc = ev.eval(x)
            G = ev.gather(G_loc)
            Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, sys_type)
            errs.append((name, abs(F - Fo) / max(1.0, abs(Fo)), float(np.max(np.abs(G - Go)) / max(np.max(np.abs(Go)), 1e-6))))
        ev.close()
    q.put((rank, errs))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_native_slice_parallel(world):
    """qoc_eval_slice over `world` GPUs: peer-memory exchange of the range propagators, boundary operators on the library's
    own GEMM kernel over NVLink; same F and G as the single-device evaluation / the oracle, on every rank."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_native_slice_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, errs in res:
        for name, ef, eg in errs:
            assert ef < 1e-10 and eg < 1e-8, (rank, name, ef, eg)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("D,M", [(8, 37), (32, 5)])
def test_threaded_multi_device_handle_on_distinct_gpus(D, M):
    """Distinct device ordinals select the threaded mode of the multi-device handle (one launch thread per device, in-process
    peer-memory all-reduce), for the warp-resident path and for the D > 16 GEMM pipeline."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import quoptimalcontrol_jl_b200 as qoc
    from oracle import grape_oracle as orc
    from conftest import assert_parity, random_system
    K, N, T = 2, 20, 0.9
    members = [random_system(D, K, seed=330 + k) for k in range(M)]
    wts = np.random.default_rng(M).random(M) / M
    devs = list(range(min(torch.cuda.device_count(), 4)))
    with qoc.GrapeEvaluator(members, T, N, orc.STATE_TRANSFER, wts=wts, devices=devs, pure_state=False, penalty=(0.01, 0.02)) as ev:
        for it in range(4):
            x = np.random.default_rng(it).uniform(-1, 1, (K, N))
            F, G = ev.eval(x)
            F0, _ = ev.eval(x, want_grad=False)
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, x, T, orc.STATE_TRANSFER)
        d = np.diff(x, axis=1)
        Fp = 0.01 * np.sum(x ** 2) + 0.02 * np.sum(d ** 2)
        Gp = 2 * 0.01 * x
        Gp[:, :-1] -= 2 * 0.02 * d
        Gp[:, 1:] += 2 * 0.02 * d
        assert_parity(F, G, Fo + Fp, Go + Gp)
        assert abs(F0 - (Fo + Fp)) < 1e-10
