"""Tuning aid: n-qubit Ising state transfer ensembles on the large-D path, chains batched vs one at a time.
`python tools/time_big_ensemble.py 5:64:500 6:32:500` = n qubits : members : slices."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quoptimalcontrol_jl_b200 as qoc
for spec in sys.argv[1:] or ["5:64:500"]:
    n, M, N = (int(v) for v in spec.split(":"))
    cfg = qoc.configs.config5(n=n, N=N)
    A, B, Xi, Xt = cfg["members"][0]
    members = [(A * (1 + 0.01 * (k - M / 2) / M), B, Xi, Xt) for k in range(M)]
    res = {}
    for batch in ("1", None):
        if batch is None: os.environ.pop("QOC_BIG_BATCH", None)
        else: os.environ["QOC_BIG_BATCH"] = batch
        with qoc.GrapeEvaluator(members, cfg["T"], N, cfg["sys_type"]) as ev:
            for _ in range(2): F, G = ev.eval(cfg["x"])
            t0 = time.perf_counter()
            for _ in range(3): ev.eval(cfg["x"])
            dt = (time.perf_counter() - t0) / 3
            res[batch] = (dt, F, G, ev.stats()["launches_last_eval"])
    (d1, F1, G1, l1), (db, Fb, Gb, lb) = res["1"], res[None]
    print(f"n={n} D={2**n} M={M} N={N}: one-at-a-time {d1*1e3:9.3f} ms ({l1} launches)  batched {db*1e3:9.3f} ms ({lb} launches)"
          f"  x{d1/db:5.2f}  per chain {db/M*1e3:7.3f} ms  dF {abs(F1-Fb):.1e} dG {np.max(np.abs(G1-Gb)):.1e}", flush=True)
