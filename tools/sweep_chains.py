"""Tuning aid: chain_kernel time vs number of chains (cfg4 members), fused path forced."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quoptimalcontrol_jl_b200 as qoc
cfg = qoc.configs.config4()
for M in [int(a) for a in sys.argv[1:]] or [148, 296, 592, 1184, 2368, 2960, 4096]:
    members = (cfg["members"] * 2)[:M]
    with qoc.GrapeEvaluator(members, cfg["T"], cfg["N"], cfg["sys_type"], wts=np.full(M, 1.0 / M)) as ev:
        for _ in range(3): ev.eval(cfg["x"])
        ev.stats()
        for _ in range(5): ev.eval(cfg["x"])
        st = ev.stats()
    ms = st["main_kernel_ms_avg"]
    print(f"M={M:5d} warps/SMSP={M/592:5.2f} kernel={ms:7.3f} ms  us/slice/warp={ms*1e3/cfg['N']:6.3f}  TFLOP/s={M*cfg['N']*4096*6/ms/1e9:6.2f}", flush=True)
