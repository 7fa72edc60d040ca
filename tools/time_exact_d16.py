"""Tuning aid: throughput of the D = 16 exact-gradient evaluation (cfg3 operators, `gradient="exact"`) for R pulses."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quoptimalcontrol_jl_b200 as qoc
cfg = qoc.configs.config3()
K, N = cfg["x"].shape
for R in [int(a) for a in sys.argv[1:]] or [1, 1024]:
    xs = np.concatenate([cfg["x"][None], np.random.default_rng(1).uniform(-1, 1, (R - 1, K, N))]) if R > 1 else cfg["x"][None]
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], N, cfg["sys_type"], gradient="exact", n_pulses=R) as ev:
        x = torch.tensor(np.ascontiguousarray(np.swapaxes(xs, 1, 2)), device="cuda")
        fg = torch.zeros((R, N * K + 1), dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream()
        for _ in range(2): ev.eval_device(x.data_ptr(), fg.data_ptr(), True, st.cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if R > 1 else 50
        e0.record(st)
        for _ in range(reps): ev.eval_device(x.data_ptr(), fg.data_ptr(), True, st.cuda_stream)
        e1.record(st); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out = fg.cpu().numpy()
        print(f"D=16 exact R={R:5d}: {ms:9.4f} ms/step  {R / ms * 1e3:10.1f} evals/s  launches {ev.stats()['launches_last_eval']}  F0={out[0,0]:.12f} |G|={np.abs(out[0,1:]).max():.6e}", flush=True)
