#!/usr/bin/env python
"""bench.py — GRAPE fidelity+gradient evaluations/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg4|cfg1|cfg2|cfg3|cfg5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...
    python bench.py --gpus N --single-process        (one process, one multi-device handle: qoc_desc.n_devices = N)

One "step" = one evaluation of the (F, G, x) closure (solve.jl:75-100 / :164-196) for the whole workload: all
ensemble members weighted and summed (and, for N > 1, one all-reduce of [F|G]).  Default workload: BASELINE.json
configs[3], the 4096-member robust ensemble (the config the 1/2/4/8-GPU metric is quoted on); its members are sharded
over the ranks (strong scaling: total work fixed).  Prints ONE JSON line on rank 0.

The parity object compares the TIMED handle's own output (the device buffer left by the last timed step, i.e. after the
all-reduce when N > 1, and the last end-to-end call's host result) with the CPU restatement over ALL members.
At N = 1 the line also carries `secondary`: the other BASELINE configs (cfg5 dense D = 256 path with full-size parity
against tests/golden/cfg5_full*.npz; cfg1-3 at one pulse and at a saturating multi-start batch).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GRAPE fidelity+gradient evals/sec"
UNIT = "evals/s"
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "fp64_peak.json")
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic.json")
GOLDEN = os.path.join(ROOT, "tests", "golden")

_OUT = sys.stdout
DEFAULT_PULSES = {"cfg4": 1, "cfg5": 1, "cfg1": 65536, "cfg2": 4096, "cfg3": 1024}


def build_config(name):
    import quoptimalcontrol_jl_b200 as qoc
    c = qoc.configs
    try:
        return {"cfg1": c.config1, "cfg2": c.config2, "cfg3": c.config3, "cfg4": c.config4, "cfg5": c.config5}[name]()
    except KeyError:
        raise SystemExit(f"unknown config {name}")


def config_dict(cfg, pulses_per_step):
    """The workload description, identical in both arms (no prose that differs between them)."""
    A, B = cfg["members"][0][0], cfg["members"][0][1]
    D, K, N, M = A.shape[0], len(B), int(cfg["N"]), len(cfg["members"])
    store_gb = M * pulses_per_step * N * D * D * 16 / 1e9
    return {"workload": cfg["name"], "D": D, "K": K, "N": N, "M": M, "pulses_per_step": int(pulses_per_step),
            "gradient": cfg["gradient"],
            "l2": "no flush: every step streams its per-slice propagator store (%.2f GB) >> 126 MB L2" % store_gb
            if store_gb > 0.5 else "no flush: working set %.3f GB, L2-resident between steps (stated, latency-bound case)" % store_gb}


_SAMPLER_SRC = r"""
import sys, time
import pynvml as n
n.nvmlInit()
h = n.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))
mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
print("ready", flush=True)
while True:
    print(time.time(), n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM), mx, n.nvmlDeviceGetPowerUsage(h) / 1000.0, int(fn(h)), flush=True)
    time.sleep(0.002)
"""


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML polled every ~2 ms by a helper
    PROCESS (a polling thread in this process contends with the launch path of a host-driven multi-device step), with
    nvidia-smi -lms as the fallback.  Started well before the timed loop so that short timed regions still get rows.
    Timestamps are time.time() in both processes."""
    BITS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.rows, self.proc, self.source = [], None, None
        try:
            self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            if self.proc.stdout.readline().strip() != "ready":
                raise OSError("nvml helper failed")
            self.source = "nvml"
        except Exception:
            try:
                q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
                    "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(index), "-lms", "20"],
                                             stdout=subprocess.PIPE, text=True)
                self.source = "nvidia-smi"
            except OSError:
                self.proc = None
        if self.proc is not None:
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()

    def _read(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            try:
                if self.source == "nvml":
                    f = line.split()
                    rs = int(f[4])
                    self.rows.append((float(f[0]), float(f[1]), float(f[2]), float(f[3]), {v for k, v in self.BITS.items() if rs & k}))
                else:
                    f = [x.strip() for x in line.split(",")]
                    self.rows.append((time.time(), float(f[0]), float(f[1]), float(f[2]),
                                      {n for n, v in zip(names, f[3:7]) if v.lower().startswith("active")}))
            except (ValueError, IndexError):
                continue

    def stop(self, t0, t1):
        """t0, t1: time.time() stamps bracketing the timed region."""
        if self.proc is None:
            return None
        self.proc.terminate()
        self.th.join(timeout=1.0)
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        rows = inside or [r for r in self.rows if t0 - 0.25 <= r[0] <= t1 + 0.05]      # nearest rows under the same load
        if not rows:
            return None
        reasons = set().union(*[r[4] for r in rows])
        return {"sm_mhz": float(np.median([r[1] for r in rows])), "sm_max_mhz": float(max(r[2] for r in rows)),
                "power_w_max": float(max(r[3] for r in rows)), "samples": len(rows), "samples_in_timed_region": len(inside),
                "source": self.source, "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------------ oracle side
def oracle_eval(cfg, threads, members=None, wts=None, x=None, N=None):
    """(F, G, seconds, description) of the CPU restatement on the given workload (default: the whole config)."""
    from oracle import c_oracle, grape_oracle
    members = cfg["members"] if members is None else members
    wts = cfg["wts"] if wts is None else wts
    x = cfg["x"] if x is None else x
    N = cfg["N"] if N is None else N
    T = cfg["T"] * N / cfg["N"]
    D = members[0][0].shape[0]
    w = wts if wts is not None else np.ones(len(members))
    t0 = time.perf_counter()
    if cfg["gradient"] == "exact":
        F, G = grape_oracle.ensemble_exact(members, w, x, T, cfg["sys_type"])
        how, cores = "numpy restatement of the ADGRAPE functional with augmented-matrix derivatives", 1
    elif D > 16:
        F, G = grape_oracle.ensemble_fom_and_gradient(members, w, x, T, cfg["sys_type"])
        how, cores = "numpy/OpenBLAS restatement of the reference loop order (3K GEMMs per slice), all BLAS threads", threads
    else:
        nt = threads if len(members) > 1 else 1
        F, G = c_oracle.eval_ensemble(members, wts, x, T, cfg["sys_type"], 0, nt)
        how = "C restatement of the reference loop order" + (", OpenMP over members (the Julia reference is serial)" if nt > 1 else ", 1 thread (the reference is serial)")
        cores = nt
    return F, G, time.perf_counter() - t0, how, cores


def parity_of(Fg, Gg, Fo, Go, checked_on, floor=1e-300):
    Gg, Go = np.asarray(Gg), np.asarray(Go)
    scale = max(float(np.max(np.abs(Go))), floor)
    out = {"fom_rel_err": float(abs(Fg - Fo) / max(1.0, abs(Fo))), "grad_rel_err_inf": float(np.max(np.abs(Gg - Go)) / scale),
           "grad_inf_norm": float(np.max(np.abs(Go))), "checked_on": checked_on, "tolerance": "1e-10 fom / 1e-8 gradient (inf-norm relative)"}
    out["ok"] = bool(out["fom_rel_err"] <= 1e-10 and out["grad_rel_err_inf"] <= 1e-8)
    return out


def run_reference(args, cfg, rank):
    """--impl reference: the reference's own CPU implementation of the path, restated in C / numpy (Julia being
    unavailable), all host threads.  Every step evaluates the WHOLE workload (all members, all slices) except for cfg5,
    where a step is a bounded sample of slices (stated)."""
    if rank != 0:
        return
    M = len(cfg["members"])
    D = cfg["members"][0][0].shape[0]
    threads = os.cpu_count() or 1
    if D > 16:
        ns = min(cfg["N"], 8)
        kw = dict(x=cfg["x"][:, :ns], N=ns)
        scale, sample = cfg["N"] / ns, f"{ns} of {cfg['N']} slices per step, scaled x{cfg['N'] / ns:g}"
    else:
        kw, scale, sample = {}, 1.0, f"all {M} members, all {cfg['N']} slices per step (no extrapolation)"
    how, cores = "", 1
    for _ in range(args.warmup):
        _, _, _, how, cores = oracle_eval(cfg, threads, **kw)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, _, _, how, cores = oracle_eval(cfg, threads, **kw)
    dt = (time.perf_counter() - t0) / args.steps * scale
    value = 1.0 / dt
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong" if M > 1 else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "impl": "reference", "config": config_dict(cfg, 1),
            "parallelism": f"host CPU, {cores} threads (rank 0 only)",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample + "; " + how},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_OUT, flush=True)


# ------------------------------------------------------------------------------------------------------------ GPU side
def fp64_peak():
    peak, src = 37.1, "fallback constant (tools/microbench/fp64_pipes.cu DMMA burst)"
    if os.path.exists(FP64_PEAK_FILE):
        with open(FP64_PEAK_FILE) as f:
            pj = json.load(f)
        peak, src = float(pj["fp64_dmma_tflops"]), pj["source"]
    return peak, src + " — builder-measured: MEASURED_PEAKS.json has no FP64 entry (64 FP64 MAC/clk/SM x 148 SM x 1.965 GHz x 2 = 37.2)"


def products_per_slice(cfg, path, K):
    import quoptimalcontrol_jl_b200 as qoc
    unitary = cfg["sys_type"] == qoc._lib.UNITARY_GATE
    credited = (3 * (1 + 2 * K) if cfg["gradient"] == "exact" else 3) + (2 if unitary else 4) + (1 if unitary else 2)
    executed = None
    if cfg["gradient"] != "exact":
        m = cfg["members"][0]
        herm = np.array_equal(m[0], m[0].conj().T) and all(np.array_equal(b, b.conj().T) for b in m[1])
        if path == 2:
            executed = 7 if herm else (4 + 1 + (2 if unitary else 4) + (1 if unitary else 2))
        elif path == 1:
            executed = 6 if herm else credited         # closed-system conjugation recursion: 3 expm + 1 total + 2 sweep
    return credited, executed


def time_device(ev, x_dev, fg_dev, stream, steps, warmup, torch):
    for _ in range(warmup):
        ev.eval_device(x_dev.data_ptr(), fg_dev.data_ptr(), True, stream.cuda_stream)
    torch.cuda.synchronize()
    ev.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        ev.eval_device(x_dev.data_ptr(), fg_dev.data_ptr(), True, stream.cuda_stream)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def secondary_config(name, R, qoc, torch, dev, local_rank, peak, threads, steps=5):
    """One of the other BASELINE configs on this GPU: device-resident and end-to-end timing, roofline, full-size parity."""
    cfg = build_config(name)
    K, N = cfg["x"].shape
    D = cfg["members"][0][0].shape[0]
    rng = np.random.default_rng(4321)
    xs = np.concatenate([cfg["x"][None], rng.uniform(-1, 1, (R - 1, K, N))]) if R > 1 else cfg["x"][None]
    out = {"config": config_dict(cfg, R)}
    stream = torch.cuda.current_stream()
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], N, cfg["sys_type"], wts=cfg["wts"], gradient=cfg["gradient"], n_pulses=R,
                            device=local_rank, pure_state=False) as ev:
        x_dev = torch.from_numpy(np.ascontiguousarray(np.swapaxes(xs, 1, 2))).to(dev)
        fg_dev = torch.zeros((R, N * K + 1), dtype=torch.float64, device=dev)
        ms = time_device(ev, x_dev, fg_dev, stream, steps, 3, torch)
        st = ev.stats()
        xin = xs if R > 1 else xs[0]
        ev.eval(xin)
        t0 = time.perf_counter()
        for _ in range(steps):
            Fg, Gg = ev.eval(xin)
        e2e_s = (time.perf_counter() - t0) / steps
        fg = fg_dev.cpu().numpy()
        credited, executed = products_per_slice(cfg, st["path"], K)
        flops = qoc.configs.alg_flops(cfg) * R
        ratio = min(1.0, executed / credited) if executed else 1.0
        ach = flops / (ms * 1e-3) / 1e12
        out.update({"ms_per_step": ms, "value": R / (ms * 1e-3), "unit": UNIT,
                    "e2e": {"value": R / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(R * N * K * 8), "d2h_bytes_per_step": int(R * (N * K + 1) * 8)},
                    "gpu_launches_per_step": int(st["launches_last_eval"]),
                    "roofline": {"bound": "tensor", "achieved": ach * ratio, "peak": peak, "unit": "TFLOP/s", "frac": ach * ratio / peak,
                                 "achieved_credited": ach, "frac_credited": ach / peak,
                                 "products_per_slice": {"credited": credited, "executed": executed}, "alg_flops_per_step": flops,
                                 "note": "whole step timed (CUDA events), all launches"}})
        # ---- full-size parity of the timed handle's own output
        if name == "cfg5":
            z = np.load(os.path.join(GOLDEN, "cfg5_full.npz"))
            Go = np.ascontiguousarray(z["G"])
            Gd = fg[0, 1:].reshape(N, K).T
            out["parity"] = parity_of(fg[0, 0], Gd, float(z["F"]), Go, "full size (2000 slices) vs tests/golden/cfg5_full.npz, device buffer of the timed handle")
            out["parity"]["e2e_result"] = parity_of(Fg, Gg, float(z["F"]), Go, "last qoc_eval result")["ok"]
            zd = np.load(os.path.join(GOLDEN, "cfg5_full_dense.npz"))
            sys.path.insert(0, GOLDEN)
            from make_golden_cfg5 import dense_states
            ev.set_states(*dense_states(D))
            Fd, Gdn = ev.eval(cfg["x"])
            out["parity_dense_states"] = parity_of(Fd, Gdn, float(zd["F"]), zd["G"], "full size, seeded dense Xi/Xt vs tests/golden/cfg5_full_dense.npz (same handle via qoc_set_states)")
        else:
            worst = None
            for r in sorted({0, R - 1}):
                Fo, Go, _, how, _ = oracle_eval(cfg, threads, x=xs[r])
                p = parity_of(fg[r, 0], fg[r, 1:].reshape(N, K).T, Fo, Go, f"full size, pulses {sorted({0, R - 1})} of {R}; {how}; device buffer of the timed handle")
                pe = parity_of(Fg[r] if R > 1 else Fg, Gg[r] if R > 1 else Gg, Fo, Go, "")
                p["e2e_result"] = pe["ok"]
                if worst is None or p["grad_rel_err_inf"] > worst["grad_rel_err_inf"]:
                    worst = p
            out["parity"] = worst
    if name == "cfg5":      # the separate pure-state vector path (not the contract arithmetic): timing + full-size parity
        with qoc.GrapeEvaluator(cfg["members"], cfg["T"], N, cfg["sys_type"], device=local_rank) as evp:
            if evp.stats()["path"] == 3:
                for _ in range(3):
                    Fp, Gp = evp.eval(cfg["x"])
                t0 = time.perf_counter()
                for _ in range(steps):
                    evp.eval(cfg["x"])
                tp = (time.perf_counter() - t0) / steps
                z = np.load(os.path.join(GOLDEN, "cfg5_full.npz"))
                out["pure_state_path"] = {"value": 1.0 / tp, "unit": UNIT, "ms_per_step": tp * 1e3, "timed": "end to end through qoc_eval (host buffers)",
                                          "gpu_launches_per_step": int(evp.stats()["launches_last_eval"]),
                                          "parity": parity_of(Fp, Gp, float(z["F"]), z["G"], "full size vs tests/golden/cfg5_full.npz"),
                                          "note": "state-vector sweep for pure-state transfers on sparse closed systems (QOC_FLAG_NO_PURE_STATE unset): "
                                                  "same F and G to rounding, O(nnz) work per slice, no FP64-roofline fraction is claimed for it"}
    return out


def run_slice(args, cfg, rank, world, local_rank, dev, torch, dist, qoc):
    """--mode slice: one large instance, slices block-partitioned over the ranks; the whole exchange + boundary-operator
    stage runs inside the library over NVLink peer memory (qoc_eval_slice).  Strong scaling: total work fixed."""
    A, B, Xi, Xt = cfg["members"][0]
    K, N = cfg["x"].shape
    D = A.shape[0]
    if len(cfg["members"]) != 1 or D <= 16:
        raise SystemExit("--mode slice needs a single large instance (cfg5)")
    ev = qoc.NativeSliceParallelEvaluator(A, B, Xi, Xt, cfg["T"], N, cfg["sys_type"], dist=dist if world > 1 else None,
                                          device=local_rank, gradient=cfg["gradient"])
    lo, hi = ev.lo, ev.hi
    xb = np.ascontiguousarray(cfg["x"][:, lo:hi].T)            # ABI layout [N_r][K]
    F, G = np.empty(1), np.empty((hi - lo, K))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        ev.local._check(ev.local._lib.qoc_eval_slice(ev.local._h, xb.ctypes.data, F.ctypes.data, G.ctypes.data))
    barrier()
    dev_ms = 0.0
    t0, w0 = time.perf_counter(), time.time()
    for _ in range(args.steps):
        ev.local._check(ev.local._lib.qoc_eval_slice(ev.local._h, xb.ctypes.data, F.ctypes.data, G.ctypes.data))
        dev_ms += ev.local.stats()["gpu_ms_last_eval"]
    barrier()
    t1, w1 = time.perf_counter(), time.time()
    clocks = sampler.stop(w0, w1) if sampler else None
    wall_ms, dev_ms = (t1 - t0) * 1e3 / args.steps, dev_ms / args.steps
    launches = ev.local.stats()["launches_last_eval"]
    if world > 1:
        t = torch.tensor([wall_ms, dev_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall_ms, dev_ms = float(t[0]), float(t[1])
        blocks = [None] * world
        dist.all_gather_object(blocks, np.ascontiguousarray(G.T))
        Gfull = np.concatenate(blocks, axis=1)
    else:
        Gfull = np.ascontiguousarray(G.T)
    if rank == 0:
        parity = None
        gpath = os.path.join(GOLDEN, "cfg5_full.npz")
        if cfg["name"].startswith("cfg5") and os.path.exists(gpath):
            z = np.load(gpath)
            parity = parity_of(float(F[0]), Gfull, float(z["F"]), z["G"], "full size vs tests/golden/cfg5_full.npz, gradient blocks of all ranks")
        peak, peak_src = fp64_peak()
        credited, executed = products_per_slice(cfg, 2, K)
        flops_rank = qoc.configs.alg_flops(cfg) / world
        ratio = min(1.0, executed / credited) if executed else 1.0
        ach = flops_rank / (dev_ms * 1e-3) / 1e12
        line = {"metric": METRIC, "value": 1e3 / dev_ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config_dict(cfg, 1),
                "parallelism": f"slice-parallel x{world}: one instance, {N} slices block-partitioned, range propagators exchanged over NVLink peer "
                               "memory, boundary operators on the library's own DMMA GEMM kernel (qoc_eval_slice)",
                "e2e": {"value": 1e3 / wall_ms, "unit": UNIT, "h2d_bytes_per_step": int((hi - lo) * K * 8), "d2h_bytes_per_step": int(((hi - lo) * K + 1) * 8),
                        "api": "qoc_eval_slice (host buffers)"},
                "gpu_launches": int(launches) * args.steps, "clocks": clocks,
                "roofline": {"bound": "tensor", "achieved": ach * ratio, "peak": peak, "unit": "TFLOP/s", "frac": ach * ratio / peak,
                             "achieved_credited": ach, "frac_credited": ach / peak, "products_per_slice": {"credited": credited, "executed": executed},
                             "traffic": None, "kernel": "zgemm_dmma_kernel x %d launches per step per rank" % launches, "kernel_ms": dev_ms,
                             "alg_flops_per_launch": flops_rank, "peak_source": peak_src},
                "cpu_baseline": None, "parity": parity,
                "timing_note": "`value`: CUDA events around the device part of qoc_eval_slice (max over ranks); `e2e`: wall clock of the same calls"}
        print(json.dumps(line), file=_OUT, flush=True)
    ev.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--pulses", type=int, default=0, help="multi-start pulses per step (0 = config default)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU restatement (no cpu_baseline, no parity)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configs (N = 1 default run only)")
    ap.add_argument("--single-process", action="store_true",
                    help="one process drives --gpus devices through ONE multi-device handle (qoc_desc.n_devices), no torchrun")
    ap.add_argument("--mode", default="ensemble", choices=["ensemble", "slice"],
                    help="slice: ONE instance (cfg5) with its time slices split over the ranks (qoc_eval_slice), strong scaling")
    ap.add_argument("--allreduce", default="oneshot", choices=["oneshot", "nccl"],
                    help="N > 1: fused one-shot all-reduce over NVLink peer memory (default) or torch.distributed NCCL")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly one JSON line: keep a private handle to it and point fd 1 at stderr, so that anything a
    # library prints to stdout (e.g. the NCCL version banner) cannot end up in front of the result
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = build_config(args.config)
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    import quoptimalcontrol_jl_b200 as qoc
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if args.mode == "slice":
        run_slice(args, cfg, rank, world, local_rank, dev, torch, dist, qoc)
        return
    single_proc = args.single_process and world == 1 and args.gpus > 1
    n_gpus = args.gpus if single_proc else world

    members, wts = cfg["members"], cfg["wts"]
    M = len(members)
    K, N = cfg["x"].shape
    D = members[0][0].shape[0]
    R = args.pulses or DEFAULT_PULSES[args.config]
    sharded = M > 1
    if sharded and not single_proc:   # ensemble members sharded over ranks, one all-reduce of [F|G] per evaluation
        lo, hi = rank * M // world, (rank + 1) * M // world
        my_members, my_wts, my_R = members[lo:hi], wts[lo:hi], R
        parallelism = f"ensemble-sharded x{world}, one process per GPU" if world > 1 else "single GPU"
    elif single_proc:
        if not sharded:
            raise SystemExit("--single-process shards ensemble members: use an ensemble config (cfg4)")
        my_members, my_wts, my_R = members, wts, R
        parallelism = f"ensemble-sharded x{n_gpus} inside ONE process (multi-device handle, qoc_desc.n_devices)"
    else:         # M = 1: replicas over independent multi-start pulses, no collective
        my_members, my_wts, my_R = members, wts, R
        parallelism = f"multi-start replicas x{world}, no collective" if world > 1 else "single GPU"
    rng = np.random.default_rng(1234 + (0 if sharded else rank))
    xs = np.concatenate([cfg["x"][None], rng.uniform(-1, 1, (my_R - 1, K, N))]) if my_R > 1 else cfg["x"][None]
    xin = xs if my_R > 1 else xs[0]

    # pure_state=False: the contract figure is the dense 9-products-per-slice evaluation; the pure-state
    # vector path changes the algorithmic FLOP count and is reported separately ("pure_state_path")
    ev = qoc.GrapeEvaluator(my_members, cfg["T"], N, cfg["sys_type"], wts=my_wts, gradient=cfg["gradient"], n_pulses=my_R,
                            device=local_rank, pure_state=False, devices=list(range(n_gpus)) if single_proc else None)
    x_host = torch.from_numpy(np.ascontiguousarray(np.swapaxes(xs, 1, 2))).pin_memory()     # [R][N][K]
    x_dev = x_host.to(dev)
    fg_dev = torch.zeros((my_R, N * K + 1), dtype=torch.float64, device=dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    oneshot = world > 1 and sharded and args.allreduce == "oneshot"
    if oneshot:                                    # exchange the CUDA IPC handles of the per-rank exchange buffers
        try:
            handles = [None] * world
            dist.all_gather_object(handles, ev.comm_export())
            ev.comm_connect(world, rank, handles)
            ok = 1
        except Exception:                          # e.g. CUDA IPC not permitted on this box
            ok = 0
        t = torch.tensor([ok], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)   # all ranks must agree on the collective they use
        if int(t.item()) == 1:
            parallelism += " + one-shot NVLink peer-memory all-reduce fused with the member reduction (graph-replayed)"
        else:
            oneshot = False
            parallelism += " + NCCL all-reduce (one-shot unavailable)"
    elif world > 1 and sharded:
        parallelism += " + NCCL all-reduce"

    last = {}
    # caller-owned host buffers in the ABI's layout (what the Julia glue passes to ccall): no per-call allocation
    xb_host = np.ascontiguousarray(np.swapaxes(xs, 1, 2))
    F_host, G_host = np.empty(my_R), np.empty((my_R, N, K))

    call_host = ev.raw_caller(xb_host, F_host, G_host)                         # qoc_eval on the caller-owned buffers
    call_host_ar = ev.raw_caller(xb_host, F_host, G_host, allreduce=True)      # qoc_eval_allreduce

    def step_device():
        if single_proc:                            # a multi-device handle has the host-buffer entry only
            call_host()
        elif oneshot:
            ev.eval_allreduce_device(x_dev.data_ptr(), fg_dev.data_ptr(), True, stream.cuda_stream)
        else:
            ev.eval_device(x_dev.data_ptr(), fg_dev.data_ptr(), True, stream.cuda_stream)
            if world > 1 and sharded:
                dist.all_reduce(fg_dev)

    def step_e2e():                                # the public host-buffer call: H2D + kernels (+ all-reduce) + D2H inside
        if oneshot:
            call_host_ar()
        elif world > 1 and sharded:
            x_dev.copy_(x_host, non_blocking=True)
            step_device()
            last["fg"] = fg_dev.cpu()
        else:
            call_host()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # untimed pre-spin of ~0.4 s: brings the clocks up and lets the sampler start.  The step count is agreed between the
    # ranks (every rank must issue the same sequence of all-reduce calls).
    t_spin = time.perf_counter()
    for _ in range(4):
        step_device()
    torch.cuda.synchronize()
    n_spin = int(min(2000, max(4, 0.4 / max((time.perf_counter() - t_spin) / 4, 1e-5))))
    if world > 1:
        t = torch.tensor([n_spin], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        n_spin = int(t.item())
    for _ in range(n_spin):
        step_device()
    barrier()
    for _ in range(args.warmup):
        step_device()
    barrier()
    ev.stats()                                    # reset the kernel-event ring
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start, w_start = time.perf_counter(), time.time()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    t_end, w_end = time.perf_counter(), time.time()
    clocks = sampler.stop(w_start, w_end) if sampler else None
    ms = (t_end - t_start) * 1e3 / args.steps if single_proc else e0.elapsed_time(e1) / args.steps
    st = ev.stats()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    evals_per_step = R if (sharded or single_proc) else R * world
    value = evals_per_step / (ms * 1e-3)
    timed_fg = None if single_proc else fg_dev.cpu().numpy()       # what the LAST TIMED step left on the device

    # ---- end-to-end timing through the public API (host buffers, copies inside the timed region) ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if oneshot or world == 1:                      # the last end-to-end result, for the parity check below
        Gk = np.swapaxes(G_host, 1, 2)
        last["F"], last["G"] = (F_host.copy(), Gk.copy()) if my_R > 1 else (float(F_host[0]), Gk[0].copy())
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": evals_per_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(my_R * N * K * 8) * (n_gpus if single_proc else 1),
           "d2h_bytes_per_step": int(my_R * (N * K + 1) * 8),
           "api": "qoc_eval_allreduce (host buffers, one CUDA-graph launch per rank)" if oneshot else "qoc_eval (host buffers)"}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    flops_total = qoc.configs.alg_flops(cfg) * R                    # algorithmic FLOPs of one step, whole job
    flops_rank = flops_total / n_gpus if sharded else flops_total   # per launch of one GPU's kernels
    peak, peak_src = fp64_peak()
    k_ms = (st["main_kernel_ms_avg"] or ms) if not single_proc else ms
    if st["path"] == 1:
        kernel_name = ("warp-resident FP64 DMMA chain kernels (chunk_expm + boundary + sweep, or chain_kernel): %d launches per "
                       "evaluation incl. the reduction; events around the chain-kernel group" % st["launches_last_eval"])
    else:
        kernel_name = "zgemm_dmma_kernel x %d launches per step (whole step timed: GEMMs are >98%% of it)" % st["launches_last_eval"]
    achieved = flops_rank / (k_ms * 1e-3) / 1e12
    traffic = None
    if os.path.exists(TRAFFIC_FILE):
        with open(TRAFFIC_FILE) as f:
            traffic = json.load(f).get(args.config)
    credited, executed = products_per_slice(cfg, st["path"], K)
    # `achieved` / `frac` count the products the kernels actually execute (never more than the credited algorithmic count)
    ratio = min(1.0, executed / credited) if executed else 1.0
    # D <= 16: every complex product runs in Gauss's three-multiplication form (3 real DMMA products + operand sums).  FLOPs
    # keep the standard 8 D^3 convention of a complex product (as for zgemm3m); the real multiply-adds actually issued are
    # reported beside it so the FP64-pipe occupancy can be read off as well.
    real_ratio = 0.75 if st["path"] == 1 else 1.0
    roofline = {"bound": "tensor", "achieved": achieved * ratio, "peak": peak, "unit": "TFLOP/s", "frac": achieved * ratio / peak,
                "achieved_credited": achieved, "frac_credited": achieved / peak,
                "products_per_slice": {"credited": credited, "executed": executed},
                "complex_product": "3M (Gauss): 6 DMMA.8x8x4 + 6 DADD per complex 8x8x8 product" if st["path"] == 1 else "4M",
                "frac_real_multiplies_issued": achieved * ratio * real_ratio / peak,
                "traffic": traffic, "kernel": kernel_name, "kernel_ms": k_ms,
                "kernel_samples": st["main_kernel_samples"], "alg_flops_per_launch": flops_rank,
                "executed_flops_per_launch": flops_rank * ratio, "peak_source": peak_src}

    # ---- CPU restatement over the WHOLE workload: parity of the timed handle's own output (+ cpu_baseline at N = 1) ----
    cpu_baseline, parity = None, None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        if D > 16:
            z = np.load(os.path.join(GOLDEN, "cfg5_full.npz")) if args.config == "cfg5" else None
            if z is not None and timed_fg is not None:
                parity = parity_of(timed_fg[0, 0], timed_fg[0, 1:].reshape(N, K).T, float(z["F"]), z["G"],
                                   "full size vs tests/golden/cfg5_full.npz, device buffer of the timed handle")
            ns = 8
            _, _, tc, how, cores = oracle_eval(cfg, threads, x=cfg["x"][:, :ns], N=ns)
            cpu_baseline = {"value": 1.0 / (tc * N / ns), "unit": UNIT, "cores": cores, "kind": "port", "sample": f"{ns} of {N} slices, {how}; scaled x{N / ns:g}"}
        else:
            Fo, Go, tc, how, cores = oracle_eval(cfg, threads)
            where = f"all {M} members x {N} slices" if M > 1 else f"full size ({N} slices)"
            if timed_fg is not None:
                parity = parity_of(timed_fg[0, 0], timed_fg[0, 1:].reshape(N, K).T, Fo, Go,
                                   f"{where}: device buffer left by the last TIMED step" + (f" (after the {world}-rank all-reduce)" if world > 1 else "") + f"; {how}")
            if "F" in last:
                Fl, Gl = (last["F"][0], last["G"][0]) if my_R > 1 else (last["F"], last["G"])
                pe = parity_of(Fl, Gl, Fo, Go, f"{where}: host result of the last end-to-end call; {how}")
                if parity is None:
                    parity = pe
                else:
                    parity["e2e_result"] = {k: pe[k] for k in ("fom_rel_err", "grad_rel_err_inf", "ok")}
            if world == 1 and not single_proc:
                cpu_baseline = {"value": 1.0 / tc, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{where}, one evaluation (no extrapolation); {how}"}

    # ---- the other BASELINE configs (N = 1, default workload only) ----
    secondary = None
    if world == 1 and not single_proc and args.config == "cfg4" and not args.no_secondary:
        ev.close()
        del x_dev, fg_dev
        torch.cuda.empty_cache()
        secondary = {}
        threads = os.cpu_count() or 1
        for name, r in (("cfg5", 1), ("cfg1", 1), ("cfg1", 65536), ("cfg2", 1), ("cfg2", 4096), ("cfg3", 1), ("cfg3", 1024)):
            key = name if r == 1 else f"{name}_batch{r}"
            try:
                secondary[key] = secondary_config(name, r, qoc, torch, dev, local_rank, peak, threads)
            except Exception as exc:   # a secondary figure must never take the headline down with it
                secondary[key] = {"error": f"{type(exc).__name__}: {exc}"}
            torch.cuda.empty_cache()

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(cfg, evals_per_step), "parallelism": parallelism,
            "e2e": e2e, "gpu_launches": int(st["launches_last_eval"]) * args.steps, "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity}
    if single_proc:
        line["timing_note"] = "single-process multi-device handle: `value` is timed through qoc_eval with host buffers (wall clock, sync per step); it has no device-pointer entry"
    if secondary is not None:
        line["secondary"] = secondary
    print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
