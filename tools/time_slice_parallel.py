"""Slice-parallel cfg5 (one D = 256 instance, dense path) over the ranks of a torchrun launch:
`python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/time_slice_parallel.py [n_qubits]`.
Prints the per-evaluation wall time (max over ranks) next to the single-device time of rank 0."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import quoptimalcontrol_jl_b200 as qoc
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = qoc.configs.config5(n=n)
A, B, Xi, Xt = cfg["members"][0]
I = np.eye(A.shape[0], dtype=complex)
unitary = cfg["sys_type"] == qoc._lib.UNITARY_GATE
ev = qoc.SliceParallelEvaluator(Xi, Xt, cfg["T"], cfg["N"], unitary,
                                lambda ns, dur: qoc.GrapeEvaluator([(A, B, I, I)], dur, ns, cfg["sys_type"], device=local, pure_state=False), dist=dist,
                                device=torch.device("cuda", local))
for _ in range(2): F, G = ev.eval(cfg["x"])
dist.barrier(); t0 = time.perf_counter()
reps = 5
for _ in range(reps): F, G = ev.eval(cfg["x"])
dt = torch.tensor([(time.perf_counter() - t0) / reps], dtype=torch.float64)
dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    with qoc.GrapeEvaluator([(A, B, Xi, Xt)], cfg["T"], cfg["N"], cfg["sys_type"], device=local, pure_state=False) as one:
        for _ in range(2): F1, G1 = one.eval(cfg["x"])
        t0 = time.perf_counter()
        for _ in range(3): one.eval(cfg["x"])
        d1 = (time.perf_counter() - t0) / 3
    print(f"n={n} D={2**n} N={cfg['N']} ranks={world}: slice-parallel {dt.item()*1e3:8.3f} ms/eval  single device {d1*1e3:8.3f} ms/eval  x{d1/dt.item():.2f}"
          f"  |dF| {abs(F-F1):.1e} max|dG| {np.max(np.abs(G-G1)):.1e} (max|G| {np.max(np.abs(G1)):.1e})", flush=True)
dist.barrier()
dist.destroy_process_group()
