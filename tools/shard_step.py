"""Tuning aid: whole-step device time (chain kernels + member reduction) of a cfg4 shard of M members on one GPU.
`python tools/shard_step.py 512 1024 4096`; `512:11` forces 11 chunks per chain (QOC_CHUNKS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import quoptimalcontrol_jl_b200 as qoc
cfg = qoc.configs.config4()
K, N = cfg["x"].shape
for spec in sys.argv[1:] or ["512"]:
    M = int(spec.split(":")[0])
    if ":" in spec: os.environ["QOC_CHUNKS"] = spec.split(":")[1]
    else: os.environ.pop("QOC_CHUNKS", None)
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
        print(f"M={spec:>8s} step {e0.elapsed_time(e1)/reps:7.4f} ms  chain kernels {s['main_kernel_ms_avg']:7.4f} ms  launches {s['launches_last_eval']}", flush=True)
