"""Tuning aid: n-qubit Ising state transfer (config5 family) timing, e.g. `python tools/time_config5.py 5 6 7`."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quoptimalcontrol_jl_b200 as qoc
for n in [int(a) for a in sys.argv[1:]] or [5, 6]:
    cfg = qoc.configs.config5(n=n)
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], cfg["N"], cfg["sys_type"]) as ev:
        for _ in range(2): ev.eval(cfg["x"])
        t0 = time.perf_counter()
        for _ in range(5): ev.eval(cfg["x"])
        dt = (time.perf_counter() - t0) / 5
    D = 2 ** n
    print(f"n={n} D={D} N={cfg['N']} {dt*1e3:8.3f} ms/eval  credited {cfg['N']*8*D**3*9/dt/1e12:6.2f} TFLOP/s", flush=True)
