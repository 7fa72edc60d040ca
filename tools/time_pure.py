"""Tuning aid: config5 family (n-qubit Ising state transfer) on the pure-state vector path vs the dense GEMM path.
`python tools/time_pure.py 8 [members]`"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quoptimalcontrol_jl_b200 as qoc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = qoc.configs.config5(n=n)
A, B, Xi, Xt = cfg["members"][0]
members = [(A * (1 + 0.01 * k / M), B, Xi, Xt) for k in range(M)]
out = {}
for pure in (True, False):
    if not pure and M > 4: continue
    with qoc.GrapeEvaluator(members, cfg["T"], cfg["N"], cfg["sys_type"], pure_state=pure) as ev:
        for _ in range(2): F, G = ev.eval(cfg["x"])
        t0 = time.perf_counter()
        for _ in range(5): ev.eval(cfg["x"])
        dt = (time.perf_counter() - t0) / 5
        st = ev.stats()
    out[pure] = (F, G)
    print(f"n={n} D={2**n} M={M} N={cfg['N']} path={st['path']} {dt*1e3:9.3f} ms/eval  gpu {st['gpu_ms_last_eval']:.3f} ms  launches {st['launches_last_eval']}", flush=True)
if False in out:
    print("F", out[True][0], out[False][0], "max|dG|", np.max(np.abs(out[True][1] - out[False][1])), "max|G|", np.max(np.abs(out[False][1])))
