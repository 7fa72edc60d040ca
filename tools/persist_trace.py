"""Tuning aid: per-item timeline of the persistent closed-system kernel (QOC_PERSIST=1, QOC_PERSIST_TRACE=<file>).
`python tools/persist_trace.py 512 gpurun_out/trace512.txt [QOC_CHUNKS=24 ...]`"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
M, path = int(sys.argv[1]), sys.argv[2]
for kv in sys.argv[3:]:
    k, v = kv.split("="); os.environ[k] = v
os.environ["QOC_PERSIST"] = "1"; os.environ["QOC_PERSIST_TRACE"] = path
import numpy as np, torch
import quoptimalcontrol_jl_b200 as qoc
cfg = qoc.configs.config4()
K, N = cfg["x"].shape
members = (cfg["members"] * 2)[:M]
with qoc.GrapeEvaluator(members, cfg["T"], N, cfg["sys_type"], wts=np.full(M, 1.0 / M)) as ev:
    x = torch.tensor(np.ascontiguousarray(cfg["x"].T), device="cuda")
    fg = torch.zeros(N * K + 1, dtype=torch.float64, device="cuda")
    for _ in range(6):
        ev.eval_device(x.data_ptr(), fg.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
print("trace written to", path)
