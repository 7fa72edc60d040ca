"""Host-side logic that needs no GPU: problem types, ensemble expansion, BASELINE config builders."""
import numpy as np

import quoptimalcontrol_jl_b200 as qoc
from oracle import grape_oracle as orc


def test_problem_types_and_old_names():
    Z = np.zeros((2, 2), dtype=complex)
    p = qoc.ClosedStateTransfer(B=[Z], A=Z, Xi=Z, Xt=Z, T=1.0, n_controls=1, guess=np.zeros((1, 3)))
    assert isinstance(p.sys_type, qoc.StateTransfer) and p.sys_type.code == orc.STATE_TRANSFER
    assert isinstance(qoc.UnitarySynthesis(B=[Z], A=Z, Xi=Z, Xt=Z, T=1.0, n_controls=1, guess=None).sys_type, qoc.UnitaryGate)
    assert qoc.OpenSystemCoherenceTransfer(B=[Z], A=Z, Xi=Z, Xt=Z, T=1.0, n_controls=1, guess=None).sys_type.code == orc.COHERENCE_TRANSFER


def test_init_ensemble_follows_reference_fixture():
    """test/setup_tests.jl:31-49: detuning sweep A_gens(k) = (k - 2.5)/2.5 * Sz * 5, alternating targets."""
    Sz = np.diag([0.5, -0.5]).astype(complex)
    r0, r1 = np.diag([1, 0]).astype(complex), np.diag([0, 1]).astype(complex)
    prob = qoc.Problem(B=[Sz], A=Sz, Xi=r0, Xt=r1, T=5.0, n_controls=1, guess=np.zeros((1, 25)), sys_type=qoc.StateTransfer())
    ens = qoc.EnsembleProblem(prob=prob, n_ens=5, A_g=lambda k: (k - 2.5) / 2.5 * Sz * 5, B_g=lambda k: [Sz],
                              XiG=lambda k: r0, XtG=lambda k: r1 if k % 2 else r0, wts=np.ones(5) / 5)
    ms = qoc.init_ensemble(ens)
    assert len(ms) == 5
    assert np.allclose(ms[0].A, -3.0 * Sz) and np.allclose(ms[4].A, 5.0 * Sz)
    assert np.allclose(ms[0].Xt, r1) and np.allclose(ms[1].Xt, r0)
    assert ms[2].T == 5.0 and prob.A is Sz       # template untouched


def test_baseline_config_shapes_and_flops():
    c = qoc.configs
    c4 = c.config4(N=4, grid=2)
    assert len(c4["members"]) == 4 and c4["members"][0][0].shape == (8, 8) and len(c4["members"][0][1]) == 6
    assert abs(sum(c4["wts"]) - 1) < 1e-15
    full = dict(c4, members=[c4["members"][0]] * 4096, N=500)
    assert c.alg_flops(full) == 4096 * 500 * 4096 * 6          # 50.3 GFLOP
    c5 = c.config5(N=2, n=8)
    assert c5["members"][0][0].shape == (256, 256) and len(c5["members"][0][1]) == 16
    assert c.alg_flops(dict(c5, N=2000)) == 2000 * 8 * 256 ** 3 * 9   # 2.416 TFLOP
    c3 = c.config3(N=3)
    A = c3["members"][0][0]
    # trace preservation of the Liouvillian convention P = exp(-i dt A): vec(I)' A = 0
    vI = np.eye(4).reshape(-1, order="F")
    assert np.max(np.abs(vI @ A)) < 1e-14
    c2 = c.config2(N=10)
    assert c2["gradient"] == "exact" and c2["members"][0][3].shape == (4, 4)


def test_shard_bounds_cover_everything():
    for M in (1, 5, 64, 4096):
        for W in (1, 2, 3, 8):
            parts = [qoc.shard_bounds(M, r, W) for r in range(W)]
            assert parts[0][0] == 0 and parts[-1][1] == M
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU restatement, tier contract (4)): exactly one JSON line on stdout with the keys the
    driver reads; nothing else may reach stdout."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "cfg1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "evals/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["workload"].startswith("cfg1")


def test_slice_parallel_evaluator_single_process():
    """SliceParallelEvaluator without a process group owns all slices: set_states receives Xi / Xt unchanged and the result
    is the local evaluator's (CPU stand-in built on the oracle; the multi-rank algebra is in tests/test_dist_gloo.py)."""
    import numpy as np
    from oracle import grape_oracle as orc
    from conftest import random_system
    import quoptimalcontrol_jl_b200 as qoc

    class Local:
        def __init__(self, A, B, n, dur):
            self.A, self.B, self.N, self.T = A, B, n, dur
        def total_propagator(self, x):
            self.x = x
            return orc.pw_evolve(self.A, self.B, x, self.T / self.N, np.eye(self.A.shape[0], dtype=complex))
        def set_states(self, Xi, Xt):
            self.Xi, self.Xt = Xi, Xt
        def eval_continue(self):
            return orc.fom_and_gradient_grape(self.A, self.B, self.x, self.T, self.Xi, self.Xt, orc.STATE_TRANSFER)

    A, B, Xi, Xt = random_system(4, 2, seed=1)
    x = np.random.default_rng(0).uniform(-1, 1, (2, 5))
    ev = qoc.SliceParallelEvaluator(Xi, Xt, 1.0, 5, False, lambda n, dur: Local(A, B, n, dur))
    assert (ev.lo, ev.hi, ev.world) == (0, 5, 1)
    F, G = ev.eval(x)
    Fo, Go = orc.fom_and_gradient_grape(A, B, x, 1.0, Xi, Xt, orc.STATE_TRANSFER)
    assert abs(F - Fo) < 1e-14 and np.max(np.abs(G - Go)) < 1e-14 and np.array_equal(ev.local.Xi, Xi)
    import pytest
    with pytest.raises(ValueError):
        qoc.SliceParallelEvaluator(Xi, Xt, 1.0, 0, False, lambda n, dur: None)


def test_bench_config_is_identical_in_both_arms():
    """bench.py: the workload description must not differ between the GPU arm and --impl reference (driver's same_config)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import quoptimalcontrol_jl_b200 as qoc
    cfg = qoc.configs.config4(N=20, grid=2)
    a, b = bench.config_dict(cfg, 1), bench.config_dict(cfg, 1)
    assert a == b and set(a) == {"workload", "D", "K", "N", "M", "pulses_per_step", "gradient", "l2"}
    assert a["D"] == 8 and a["K"] == 6 and a["M"] == 4 and "parallelism" not in a
    p = bench.parity_of(1.0, np.ones((2, 3)), 1.0 + 1e-12, np.ones((2, 3)) * (1 + 1e-9), "x")
    assert p["ok"] and p["grad_rel_err_inf"] < 1.1e-9
    assert not bench.parity_of(1.0, np.ones((2, 3)), 1.0, np.ones((2, 3)) * 1.001, "x")["ok"]
