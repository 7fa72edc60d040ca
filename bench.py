#!/usr/bin/env python
"""bench.py — GRAPE fidelity+gradient evaluations/sec on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg4|cfg1|cfg2|cfg3|cfg5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one evaluation of the (F, G, x) closure (solve.jl:75-100 / :164-196) for the whole workload: all
ensemble members weighted and summed (and, for N > 1, one all-reduce of [F|G] over NCCL).  Default workload:
BASELINE.json configs[3], the 4096-member robust ensemble (the config the 1/2/4/8-GPU metric is quoted on); its
members are sharded over the ranks (strong scaling: total work fixed).  Prints ONE JSON line on rank 0.
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


_OUT = sys.stdout


def build_config(name):
    import quoptimalcontrol_jl_b200 as qoc
    c = qoc.configs
    if name == "cfg4":
        return c.config4()
    if name == "cfg1":
        return c.config1()
    if name == "cfg2":
        return c.config2()
    if name == "cfg3":
        return c.config3()
    if name == "cfg5":
        return c.config5()
    raise SystemExit(f"unknown config {name}")


DEFAULT_PULSES = {"cfg4": 1, "cfg5": 1, "cfg1": 65536, "cfg2": 4096, "cfg3": 1024}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [r for (t, r) in self.rows if (t0 is None or t >= t0) and (t1 is None or t <= t1)] or [r for _, r in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [s.strip() for s in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, cfg, rank):
    """--impl reference: the reference's own CPU implementation of the path, restated in C (oracle/grape_oracle.c,
    Julia being unavailable), all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import c_oracle, grape_oracle
    members, wts = cfg["members"], cfg["wts"]
    M = len(members)
    D = members[0][0].shape[0]
    threads = os.cpu_count() or 1
    if D > 16:
        kind_fn = "numpy"
    else:
        kind_fn = "c"
    # bounded sample: members for ensembles, slices for single long chains (cost is linear in both)
    if M > 1:
        ms = min(M, max(threads, 64))
        sample_members, sample_w, x, N = members[:ms], (wts[:ms] if wts is not None else None), cfg["x"], cfg["N"]
        scale = M / ms
        sample = f"{ms} of {M} members (all {cfg['N']} slices), scaled x{scale:g}"
    else:
        ns = min(cfg["N"], 200 if D <= 16 else 8)
        sample_members, sample_w, x, N = members, wts, cfg["x"][:, :ns], ns
        scale = cfg["N"] / ns
        sample = f"{ns} of {cfg['N']} slices, scaled x{scale:g}"
    T = cfg["T"] * N / cfg["N"]

    def step():
        if kind_fn == "c":
            if cfg["gradient"] == "exact":
                return grape_oracle.ensemble_exact(sample_members, sample_w if sample_w is not None else [1.0], x, T, cfg["sys_type"])
            return c_oracle.eval_ensemble(sample_members, sample_w, x, T, cfg["sys_type"], 0, threads)
        return grape_oracle.ensemble_fom_and_gradient(sample_members, sample_w if sample_w is not None else [1.0], x, T, cfg["sys_type"])

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = 1.0 / (dt * scale)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * scale * 1e3, "higher_is_better": True,
            "scaling": "strong" if M > 1 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": cfg["name"], "D": D, "K": int(cfg["x"].shape[0]), "N": int(cfg["N"]), "M": M, "pulses_per_step": 1,
                       "gradient": cfg["gradient"], "parallelism": f"host CPU, {threads} threads (rank 0 only)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads if kind_fn == "c" and M > 1 else (threads if kind_fn == "numpy" else 1),
                             "kind": "port", "sample": sample + ("; C restatement + OpenMP over members" if kind_fn == "c" else "; numpy/OpenBLAS restatement")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_OUT, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--pulses", type=int, default=0, help="multi-start pulses per step (0 = config default)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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

    members, wts = cfg["members"], cfg["wts"]
    M = len(members)
    K, N = cfg["x"].shape
    R = args.pulses or DEFAULT_PULSES[args.config]
    sharded = M > 1
    if sharded:   # ensemble members sharded over ranks, one all-reduce of [F|G] per evaluation
        lo, hi = rank * M // world, (rank + 1) * M // world
        my_members, my_wts, my_R = members[lo:hi], wts[lo:hi], R
        parallelism = f"ensemble-sharded x{world} + allreduce" if world > 1 else "single GPU"
    else:         # M = 1: replicas over independent multi-start pulses, no collective
        my_members, my_wts, my_R = members, wts, R
        parallelism = f"multi-start replicas x{world}, no collective" if world > 1 else "single GPU"
    rng = np.random.default_rng(1234 + (0 if sharded else rank))
    xs = np.concatenate([cfg["x"][None], rng.uniform(-1, 1, (my_R - 1, K, N))]) if my_R > 1 else cfg["x"][None]

    # pure_state=False: the contract figure is the dense 9-products-per-slice evaluation; the pure-state
    # vector path changes the algorithmic FLOP count and is reported separately below ("pure_state_path")
    ev = qoc.GrapeEvaluator(my_members, cfg["T"], N, cfg["sys_type"], wts=my_wts, gradient=cfg["gradient"],
                            n_pulses=my_R, device=local_rank, pure_state=False)
    x_host = torch.from_numpy(np.ascontiguousarray(np.swapaxes(xs, 1, 2))).pin_memory()     # [R][N][K]
    x_dev = x_host.to(dev)
    fg_dev = torch.zeros((my_R, N * K + 1), dtype=torch.float64, device=dev)
    fg_host = torch.zeros((my_R, N * K + 1), dtype=torch.float64).pin_memory()
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    oneshot = world > 1 and sharded and args.allreduce == "oneshot"
    if oneshot:                                    # exchange the CUDA IPC handles of the per-rank exchange buffers
        try:
            handles = [None] * world
            dist.all_gather_object(handles, ev.comm_export())
            ev.comm_connect(world, rank, handles)
            ok = 1
        except Exception as exc:                   # e.g. CUDA IPC not permitted on this box
            ok, why = 0, str(exc)
        t = torch.tensor([ok], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)   # all ranks must agree on the collective they use
        if int(t.item()) == 1:
            parallelism += " (one-shot NVLink peer-memory all-reduce kernel)"
        else:
            oneshot = False
            parallelism += " (NCCL all-reduce; one-shot unavailable)"
    elif world > 1 and sharded:
        parallelism += " (NCCL all-reduce)"

    def step_device():
        if oneshot:
            ev.eval_allreduce_device(x_dev.data_ptr(), fg_dev.data_ptr(), True, stream.cuda_stream)
            return
        ev.eval_device(x_dev.data_ptr(), fg_dev.data_ptr(), True, stream.cuda_stream)
        if world > 1 and sharded:
            dist.all_reduce(fg_dev)

    def step_e2e():
        if world == 1:
            return ev.eval(xs if my_R > 1 else xs[0])            # C-ABI call with host buffers (H2D + kernels + D2H)
        x_dev.copy_(x_host, non_blocking=True)
        step_device()
        fg_host.copy_(fg_dev, non_blocking=True)
        torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----
    sampler = ClockSampler(local_rank) if rank == 0 else None     # started early: nvidia-smi needs ~0.1 s to come up
    for _ in range(args.warmup):
        step_device()
    barrier()
    ev.stats()                                    # reset the kernel-event ring
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    t_end = time.perf_counter()
    clocks = sampler.stop(t_start, t_end) if sampler else None
    ms = e0.elapsed_time(e1) / args.steps
    st = ev.stats()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    evals_per_step = R if sharded else R * world
    value = evals_per_step / (ms * 1e-3)

    # ---- end-to-end timing through the public API (host buffers, copies inside the timed region) ----
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e = {"value": evals_per_step / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(my_R * N * K * 8),
           "d2h_bytes_per_step": int(my_R * (N * K + 1) * 8)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    flops_total = qoc.configs.alg_flops(cfg) * R                    # algorithmic FLOPs of one step, whole job
    flops_rank = flops_total / world if sharded else flops_total    # per launch of this rank's kernel
    peak, peak_src = 37.1, "fallback constant (tools/microbench/fp64_pipes.cu DMMA burst)"
    if os.path.exists(FP64_PEAK_FILE):
        with open(FP64_PEAK_FILE) as f:
            pj = json.load(f)
        peak, peak_src = float(pj["fp64_dmma_tflops"]), pj["source"]
    k_ms = st["main_kernel_ms_avg"] or ms
    if st["path"] == 1:
        kernel_name = ("warp-resident FP64 DMMA chain kernels (chunk_expm + boundary + sweep, or chain_kernel): %d launches per "
                       "evaluation incl. the two reduce passes; events around the chain-kernel group" % st["launches_last_eval"])
    else:
        kernel_name = "zgemm_dmma_kernel x %d launches per step (whole step timed: GEMMs are >98%% of it)" % st["launches_last_eval"]
    achieved = flops_rank / (k_ms * 1e-3) / 1e12
    traffic = None
    if os.path.exists(TRAFFIC_FILE):
        with open(TRAFFIC_FILE) as f:
            traffic = json.load(f).get(args.config)
    # products actually executed per slice (the closed-system conjugation recursion needs fewer than the credited count)
    unitary_sys = cfg["sys_type"] == qoc._lib.UNITARY_GATE
    credited = (3 * (1 + 2 * K) if cfg["gradient"] == "exact" else 3) + (2 if unitary_sys else 4) + (1 if unitary_sys else 2)
    executed = None
    if st["path"] == 2 and cfg["gradient"] != "exact":
        herm = all(np.array_equal(m[0], m[0].conj().T) and all(np.array_equal(b, b.conj().T) for b in m[1]) for m in members[:1])
        executed = 7 if herm else (4 + 1 + (2 if unitary_sys else 4) + (1 if unitary_sys else 2))
    elif st["path"] == 1 and cfg["gradient"] != "exact":
        executed = credited
    # `achieved` / `frac` count the products the kernels actually execute (never more than the credited algorithmic count):
    # the closed-system recursion needs 7 of the 9 credited products per slice, and crediting 9 would put cfg5 above the
    # pipe's peak.  The credited convention is kept alongside as achieved_credited / frac_credited.
    ratio = min(1.0, executed / credited) if executed else 1.0
    roofline = {"bound": "tensor", "achieved": achieved * ratio, "peak": peak, "unit": "TFLOP/s", "frac": achieved * ratio / peak,
                "achieved_credited": achieved, "frac_credited": achieved / peak,
                "products_per_slice": {"credited": credited, "executed": executed},
                "traffic": traffic, "kernel": kernel_name, "kernel_ms": k_ms,
                "kernel_samples": st["main_kernel_samples"], "alg_flops_per_launch": flops_rank,
                "executed_flops_per_launch": flops_rank * ratio,
                "peak_source": peak_src + " — MEASURED_PEAKS.json has no FP64 entry"}

    # ---- CPU baseline (reference restated in C, all host threads, bounded sample) + parity gate ----
    cpu_baseline, parity = None, None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import c_oracle, grape_oracle
        threads = os.cpu_count() or 1
        D = members[0][0].shape[0]
        if M > 1:  # noqa
            ms_n = min(M, max(2 * threads, 64))
            sm, sw = members[:ms_n], wts[:ms_n]
            t0 = time.perf_counter()
            Fo, Go = c_oracle.eval_ensemble(sm, sw, cfg["x"], cfg["T"], cfg["sys_type"], 0, threads)
            tc = time.perf_counter() - t0
            cpu_baseline = {"value": 1.0 / (tc * M / ms_n), "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"{ms_n} of {M} members, all {N} slices, C restatement of the reference loop order with OpenMP over members (the Julia reference is serial); scaled x{M / ms_n:g}"}
            # the parity sample has few chains: force the execution strategy of the timed run (fused, one warp per chain)
            saved = {k: os.environ.get(k) for k in ("QOC_PHASED", "QOC_CHUNKED")}
            os.environ["QOC_PHASED"] = "0"; os.environ["QOC_CHUNKED"] = "0"
            try:
                with qoc.GrapeEvaluator(sm, cfg["T"], N, cfg["sys_type"], wts=sw, gradient=cfg["gradient"], device=local_rank) as ev2:
                    Fg, Gg = ev2.eval(cfg["x"])
            finally:
                for k, v in saved.items():
                    if v is None:
                        os.environ.pop(k, None)
                    else:
                        os.environ[k] = v
        else:
            ns = min(N, 100 if D <= 16 else 16)
            xsmp = cfg["x"][:, :ns]
            Ts = cfg["T"] * ns / N
            t0 = time.perf_counter()
            pm = members
            if D > 16:
                # a 16-slice prefix cannot move |0..0> towards |1..1>: F = 1 and G ~ 1e-29 would make the parity figure
                # pure rounding noise.  The sample keeps operators, pulse and dt but uses seeded dense states.
                prng = np.random.default_rng(99)
                def dens():
                    Z = prng.standard_normal((D, D)) + 1j * prng.standard_normal((D, D))
                    Z = Z @ Z.conj().T
                    return Z / np.trace(Z).real
                pm = [(members[0][0], members[0][1], dens(), dens())]
                Fo, Go = grape_oracle.fom_and_gradient_grape(*pm[0][:2], xsmp, Ts, *pm[0][2:], cfg["sys_type"])
                how = "numpy/OpenBLAS restatement of the reference loop order (3K GEMMs per slice), all BLAS threads; parity on seeded dense random initial/target states"
            elif cfg["gradient"] == "exact":
                Fo, Go = grape_oracle.exact_fom_and_gradient(*members[0][:2], xsmp, Ts, *members[0][2:], cfg["sys_type"])
                how = "numpy restatement of the ADGRAPE functional with augmented-matrix derivatives"
            else:
                Fo, Go = c_oracle.eval_ensemble(members, None, xsmp, Ts, cfg["sys_type"], 0, 1)
                how = "C restatement of the reference loop order, 1 thread (the reference is serial)"
            tc = time.perf_counter() - t0
            cpu_baseline = {"value": 1.0 / (tc * N / ns), "unit": UNIT, "cores": threads if D > 16 else 1, "kind": "port",
                            "sample": f"{ns} of {N} slices, {how}; scaled x{N / ns:g}"}
            with qoc.GrapeEvaluator(pm, Ts, ns, cfg["sys_type"], gradient=cfg["gradient"], device=local_rank) as ev2:
                Fg, Gg = ev2.eval(xsmp)
        parity = {"fom_rel_err": abs(Fg - Fo) / max(1.0, abs(Fo)),
                  "grad_rel_err_inf": float(np.max(np.abs(Gg - Go)) / max(np.max(np.abs(Go)), 1e-6)),
                  "grad_inf_norm": float(np.max(np.abs(Go))),
                  "checked_on": cpu_baseline["sample"].split(",")[0], "tolerance": "1e-10 fom / 1e-8 gradient"}

    # ---- separate figure: pure-state vector path (same F, G; O(nnz) per slice; not the contract arithmetic) ----
    pure = None
    D = members[0][0].shape[0]
    if world == 1 and D > 16 and cfg["gradient"] != "exact":
        with qoc.GrapeEvaluator(my_members, cfg["T"], N, cfg["sys_type"], wts=my_wts, n_pulses=my_R, device=local_rank) as evp:
            if evp.stats()["path"] == 3:
                xin = xs if my_R > 1 else xs[0]
                for _ in range(3):
                    evp.eval(xin)
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    evp.eval(xin)
                tp = (time.perf_counter() - t0) / args.steps
                pure = {"value": R / tp, "unit": UNIT, "ms_per_step": tp * 1e3, "timed": "end to end through qoc_eval (host buffers)",
                        "gpu_launches_per_step": int(evp.stats()["launches_last_eval"]),
                        "note": "state-vector sweep for pure-state transfers on sparse closed systems (QOC_FLAG_NO_PURE_STATE unset): "
                                "same F and G to rounding, O(nnz) work per slice, so no FP64-roofline fraction is claimed for it"}
        if pure is not None and not args.no_cpu_baseline:
            from oracle import grape_oracle
            ns = min(N, 16)
            prng = np.random.default_rng(98)
            def ket():
                v = prng.standard_normal(D) + 1j * prng.standard_normal(D)
                v /= np.linalg.norm(v)
                return np.outer(v, v.conj())
            pmem = (members[0][0], members[0][1], ket(), ket())
            Fo, Go = grape_oracle.fom_and_gradient_grape(*pmem[:2], cfg["x"][:, :ns], cfg["T"] * ns / N, *pmem[2:], cfg["sys_type"])
            with qoc.GrapeEvaluator([pmem], cfg["T"] * ns / N, ns, cfg["sys_type"], device=local_rank) as evq:
                Fg, Gg = evq.eval(cfg["x"][:, :ns])
                assert evq.stats()["path"] == 3
            pure["parity"] = {"fom_rel_err": abs(Fg - Fo) / max(1.0, abs(Fo)),
                              "grad_rel_err_inf": float(np.max(np.abs(Gg - Go)) / max(np.max(np.abs(Go)), 1e-6)),
                              "grad_inf_norm": float(np.max(np.abs(Go))), "checked_on": f"{ns} of {N} slices, seeded random pure states"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": cfg["name"], "D": D, "K": K, "N": N, "M": M, "pulses_per_step": evals_per_step,
                       "gradient": cfg["gradient"], "parallelism": parallelism,
                       "l2": "no flush: every step writes and re-reads its per-slice propagator (and state) stores, %.2f GB of workspace >> 126 MB L2" % (st["workspace_bytes"] / 1e9)},
            "e2e": e2e, "gpu_launches": int(st["launches_last_eval"]) * args.steps, "clocks": clocks,
            "roofline": roofline, "cpu_baseline": cpu_baseline, "parity": parity}
    if pure is not None:
        line["pure_state_path"] = pure
    print(json.dumps(line), file=_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
