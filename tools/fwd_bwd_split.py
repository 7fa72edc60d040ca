"""Tuning aid: value-only (forward sweep, nothing stored) vs full fidelity+gradient time on cfg4."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import quoptimalcontrol_jl_b200 as qoc
cfg = qoc.configs.config4()
with qoc.GrapeEvaluator(cfg["members"], cfg["T"], cfg["N"], cfg["sys_type"], wts=cfg["wts"]) as ev:
    for want in (False, True):
        for _ in range(3): ev.eval(cfg["x"], want_grad=want)
        ev.stats()
        for _ in range(10): ev.eval(cfg["x"], want_grad=want)
        print("want_grad=%s kernel %.3f ms" % (want, ev.stats()["main_kernel_ms_avg"]), flush=True)
