#!/bin/bash
# Multi-GPU measurement set on one box: bash tools/run_multigpu.sh "8 4" r02x   (counts, tag)
counts=${1:-"2"}; tag=${2:-r02}
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    p=l.get("parity") or {}
    print(sys.argv[1].split("/")[-1], "n=%d value=%.1f ms=%.4f e2e=%.1f parity=%s %.1e frac=%.3f"%(l["n_gpus"], l["value"], l["ms_per_step"], l["e2e"]["value"], p.get("ok"), p.get("grad_rel_err_inf") or -1, l["roofline"]["frac"]))
except Exception as e: print(sys.argv[1], "FAILED", e)
PY
}
for n in $counts; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/${tag}_scale_n$n.json 2> gpurun_out/${tag}_scale_n$n.err; summ gpurun_out/${tag}_scale_n$n.json
  timeout 200 python bench.py --gpus $n --single-process --steps 20 --warmup 5 > gpurun_out/${tag}_single_n$n.json 2> gpurun_out/${tag}_single_n$n.err; summ gpurun_out/${tag}_single_n$n.json
  if [ -z "$NO_SLICE" ]; then timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2962$n bench.py --gpus $n --config cfg5 --mode slice --steps 5 --warmup 3 > gpurun_out/${tag}_slice_n$n.json 2> gpurun_out/${tag}_slice_n$n.err; summ gpurun_out/${tag}_slice_n$n.json; fi
done
timeout 100 python bench.py --steps 20 --warmup 5 --no-secondary > gpurun_out/${tag}_scale_n1.json 2>/dev/null; summ gpurun_out/${tag}_scale_n1.json
