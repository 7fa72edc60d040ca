"""Tuning aid: whole-step device time (chain kernels + member reduction) of a cfg4 shard of M members on one GPU.
`python tools/shard_step.py 512 1024 4096`; `512:QOC_BAL=0,QOC_CHUNKS=11` sets environment overrides for that run.
Every run is checked against the full-size result of the default strategy (max relative gradient deviation printed)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quoptimalcontrol_jl_b200 as qoc
cfg = qoc.configs.config4()
K, N = cfg["x"].shape
ref = {}
for spec in sys.argv[1:] or ["512"]:
    M = int(spec.split(":")[0])
    envs = dict(kv.split("=") for kv in spec.split(":")[1].split(",")) if ":" in spec else {}
    for k, v in envs.items(): os.environ[k] = v
    members = (cfg["members"] * 2)[:M]
    with qoc.GrapeEvaluator(members, cfg["T"], N, cfg["sys_type"], wts=np.full(M, 1.0 / M)) as ev:
        x = torch.tensor(np.ascontiguousarray(cfg["x"].T), device="cuda")
        fg = torch.zeros(N * K + 1, dtype=torch.float64, device="cuda")
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(5): ev.eval_device(x.data_ptr(), fg.data_ptr(), stream=st.cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            st.synchronize(); ev.stats()
            reps = 50
            e0.record(st)
            for _ in range(reps): ev.eval_device(x.data_ptr(), fg.data_ptr(), stream=st.cuda_stream)
            e1.record(st); st.synchronize()
        s = ev.stats()
        out = fg.cpu().numpy()
        import time
        xb = np.ascontiguousarray(cfg["x"].T); Fh = np.empty(1); Gh = np.empty((N, K))
        call = ev.raw_caller(xb, Fh, Gh)
        for _ in range(5): call()
        t0 = time.perf_counter()
        for _ in range(200): call()
        e2e_ms = (time.perf_counter() - t0) / 200 * 1e3
        if M not in ref: ref[M] = out
        dev = float(np.max(np.abs(out[1:] - ref[M][1:])) / np.max(np.abs(ref[M][1:])))
        print(f"M={spec:>40s} step {e0.elapsed_time(e1)/reps:7.4f} ms  chain kernels {s['main_kernel_ms_avg']:7.4f} ms  launches {s['launches_last_eval']}  dev vs first {dev:.1e}  host-buffer call {e2e_ms:7.4f} ms (+{(e2e_ms - e0.elapsed_time(e1)/reps)*1e3:5.1f} us)", flush=True)
    for k in envs: os.environ.pop(k, None)
