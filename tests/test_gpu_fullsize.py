"""Full-size BASELINE configs through the C ABI against the CPU oracle / the committed full-size fixtures.

The reduced-size parity tests never reach the execution strategies the full sizes take (phased pipeline with 31-44
chunks at N = 1000 / 2000, chunk-parallel closed-system kernels with 8 chunks per chain at 4096 chains, Kogge-Stone
chunk-boundary products at D = 256 with 37-74 chunks); these do.  Tolerances: north_star's (1e-10 / 1e-8)."""
import os
import sys

import numpy as np
import pytest

import quoptimalcontrol_jl_b200 as qoc
from oracle import c_oracle, grape_oracle as orc
from conftest import assert_parity

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def assert_rel(F, G, Fo, Go, ftol=1e-10, gtol=1e-8):
    """Purely relative gradient check (no absolute floor): for instances whose gradient is tiny but well conditioned."""
    assert abs(F - Fo) <= ftol * max(1.0, abs(Fo)), f"fom {F!r} vs {Fo!r}"
    scale = np.max(np.abs(Go))
    err = np.max(np.abs(np.asarray(G) - Go))
    assert err <= gtol * scale, f"gradient max err {err:.3e} vs |G|_inf {scale:.3e}"


@pytest.mark.parametrize("variant", ["test", "readme"])
def test_cfg1_full(variant):
    cfg = qoc.configs.config1(variant=variant)
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], cfg["N"], cfg["sys_type"]) as ev:
        F, G = ev.eval(cfg["x"])
    Fo, Go = orc.fom_and_gradient_grape(*cfg["members"][0][:2], cfg["x"], cfg["T"], *cfg["members"][0][2:], cfg["sys_type"])
    assert_parity(F, G, Fo, Go)


def test_cfg2_full_exact():
    """Two-qubit CNOT, N = 1000, exact gradient: single chain -> phased pipeline with ~31 chunks."""
    cfg = qoc.configs.config2()
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], cfg["N"], cfg["sys_type"], gradient="exact") as ev:
        F, G = ev.eval(cfg["x"])
        assert ev.stats()["launches_last_eval"] >= 5          # the slice-parallel pipeline, not the fused kernel
    Fo, Go = orc.exact_fom_and_gradient(*cfg["members"][0][:2], cfg["x"], cfg["T"], *cfg["members"][0][2:], cfg["sys_type"])
    assert_parity(F, G, Fo, Go)


def test_cfg2_full_first_order_both_signs():
    cfg = qoc.configs.config2()
    for conv, var in (("inplace", orc.REF_INPLACE), ("static", orc.REF_STATIC)):
        with qoc.GrapeEvaluator(cfg["members"], cfg["T"], cfg["N"], cfg["sys_type"], convention=conv) as ev:
            F, G = ev.eval(cfg["x"])
        Fo, Go = c_oracle.eval_ensemble(cfg["members"], None, cfg["x"], cfg["T"], cfg["sys_type"], var, 1)
        assert_parity(F, G, Fo, Go)


def test_cfg3_full():
    """Two-qubit Liouvillian (D = 16, non-Hermitian generator), N = 2000: phased pipeline with ~44 chunks, general kernels."""
    cfg = qoc.configs.config3()
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], cfg["N"], cfg["sys_type"]) as ev:
        F, G = ev.eval(cfg["x"])
    Fo, Go = c_oracle.eval_ensemble(cfg["members"], None, cfg["x"], cfg["T"], cfg["sys_type"], 0, 1)
    assert_parity(F, G, Fo, Go)


def test_cfg4_full_timed_strategy():
    """All 4096 members x 500 slices with the strategy bench.py times (chunk-parallel closed-system kernels, 8 chunks per
    chain), host-buffer and device-pointer entry points."""
    import torch
    cfg = qoc.configs.config4()
    threads = os.cpu_count() or 1
    Fo, Go = c_oracle.eval_ensemble(cfg["members"], cfg["wts"], cfg["x"], cfg["T"], cfg["sys_type"], 0, threads)
    K, N = cfg["x"].shape
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], N, cfg["sys_type"], wts=cfg["wts"]) as ev:
        F, G = ev.eval(cfg["x"])
        assert ev.stats()["launches_last_eval"] >= 3          # closed_persistent_kernel + reduce pass 1 + 2 (three-launch form: 3 per chain range + 2)
        assert_parity(F, G, Fo, Go)
        F0, _ = ev.eval(cfg["x"], want_grad=False)
        assert_parity(F0, None, Fo, None)
        x_dev = torch.from_numpy(np.ascontiguousarray(cfg["x"].T)).cuda()
        fg = torch.zeros(N * K + 1, dtype=torch.float64, device="cuda")
        ev.eval_device(x_dev.data_ptr(), fg.data_ptr(), True, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        fgh = fg.cpu().numpy()
        assert_parity(fgh[0], fgh[1:].reshape(N, K).T, Fo, Go)


@pytest.mark.parametrize("shard", [512, 1024])
def test_cfg4_shard_sizes(shard):
    """The per-GPU shard sizes of the 8- and 4-GPU runs (chunk-parallel mode with other chunk counts)."""
    cfg = qoc.configs.config4()
    mem, w = cfg["members"][:shard], cfg["wts"][:shard]
    Fo, Go = c_oracle.eval_ensemble(mem, w, cfg["x"], cfg["T"], cfg["sys_type"], 0, os.cpu_count() or 1)
    with qoc.GrapeEvaluator(mem, cfg["T"], cfg["N"], cfg["sys_type"], wts=w) as ev:
        F, G = ev.eval(cfg["x"])
    assert_parity(F, G, Fo, Go)


def _cfg5_golden(name):
    z = np.load(os.path.join(GOLDEN, name))
    cfg = qoc.configs.config5()
    assert np.array_equal(z["x"], cfg["x"]), "the fixture was generated for another pulse"
    return cfg, float(z["F"]), z["G"]


def test_cfg5_full_dense_path_vs_golden():
    """The literal 2000-slice |0..0> -> |1..1> instance on the tiled-GEMM path (closed-system recursion, prefix boundaries)
    against tests/golden/cfg5_full.npz (reference-order oracle run, tests/golden/make_golden_cfg5.py), then the same handle
    with seeded dense states (qoc_set_states) against cfg5_full_dense.npz."""
    cfg, Fo, Go = _cfg5_golden("cfg5_full.npz")
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], cfg["N"], cfg["sys_type"], pure_state=False) as ev:
        F, G = ev.eval(cfg["x"])
        assert ev.stats()["path"] == 2
        assert_rel(F, G, Fo, Go)
        sys.path.insert(0, GOLDEN)
        from make_golden_cfg5 import dense_states
        _, Fd, Gd = _cfg5_golden("cfg5_full_dense.npz")
        ev.set_states(*dense_states(256))
        F, G = ev.eval(cfg["x"])
        assert_rel(F, G, Fd, Gd)


def test_cfg5_full_general_path_vs_golden(monkeypatch):
    """Same instance with the closed-system shortcut disabled: separate state / costate sweeps (11 products per slice)."""
    monkeypatch.setenv("QOC_BIG_HERM", "0")
    cfg, Fo, Go = _cfg5_golden("cfg5_full.npz")
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], cfg["N"], cfg["sys_type"], pure_state=False) as ev:
        F, G = ev.eval(cfg["x"])
    assert_rel(F, G, Fo, Go)


def test_cfg5_full_pure_state_path_vs_golden():
    cfg, Fo, Go = _cfg5_golden("cfg5_full.npz")
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], cfg["N"], cfg["sys_type"]) as ev:
        F, G = ev.eval(cfg["x"])
        assert ev.stats()["path"] == 3
    assert_rel(F, G, Fo, Go)


def test_second_set_system_drops_captured_graph():
    """A handle that evaluated a Hermitian system and is then given a non-Hermitian one (same shape) must not replay the
    closed-system graph (advisor finding, round 1)."""
    from conftest import random_system
    D, K, N, T = 8, 3, 40, 1.2
    x = np.random.default_rng(0).uniform(-1, 1, (K, N))
    herm = random_system(D, K, seed=1, hermitian=True)
    nonh = random_system(D, K, seed=2, hermitian=False)
    with qoc.GrapeEvaluator([herm], T, N, orc.STATE_TRANSFER) as ev:
        for _ in range(2):
            F, G = ev.eval(x)
        assert_parity(F, G, *orc.fom_and_gradient_grape(herm[0], herm[1], x, T, herm[2], herm[3], orc.STATE_TRANSFER))
        cm = lambda m: np.ascontiguousarray(np.swapaxes(np.asarray(m, dtype=complex), -1, -2))
        a, b, xi, xt = cm(nonh[0]), cm(nonh[1]), cm(nonh[2]), cm(nonh[3])
        ev._check(ev._lib.qoc_set_system(ev._h, a.ctypes.data, b.ctypes.data, xi.ctypes.data, xt.ctypes.data, None, 0))
        F, G = ev.eval(x)
        assert_parity(F, G, *orc.fom_and_gradient_grape(nonh[0], nonh[1], x, T, nonh[2], nonh[3], orc.STATE_TRANSFER))
        a, b, xi, xt = cm(herm[0]), cm(herm[1]), cm(herm[2]), cm(herm[3])          # keep the buffers alive across the call
        ev._check(ev._lib.qoc_set_system(ev._h, a.ctypes.data, b.ctypes.data, xi.ctypes.data, xt.ctypes.data, None, 0))
        F, G = ev.eval(x)
        assert_parity(F, G, *orc.fom_and_gradient_grape(herm[0], herm[1], x, T, herm[2], herm[3], orc.STATE_TRANSFER))


def test_value_only_device_rows_are_zero_filled():
    """qoc_eval_device with want_gradient = 0 writes zeros into the G columns on both size regimes."""
    import torch
    from conftest import random_system
    for D in (4, 32):
        K, N = 2, 6
        sysm = random_system(D, K, seed=3)
        x = np.random.default_rng(1).uniform(-1, 1, (K, N))
        with qoc.GrapeEvaluator([sysm], 1.0, N, orc.STATE_TRANSFER, pure_state=False) as ev:
            x_dev = torch.from_numpy(np.ascontiguousarray(x.T)).cuda()
            fg = torch.full((N * K + 1,), 7.0, dtype=torch.float64, device="cuda")
            ev.eval_device(x_dev.data_ptr(), fg.data_ptr(), True, None)
            torch.cuda.synchronize()
            assert float(fg[1:].abs().max()) > 0
            ev.eval_device(x_dev.data_ptr(), fg.data_ptr(), False, None)
            torch.cuda.synchronize()
            Fo, _ = orc.fom_and_gradient_grape(sysm[0], sysm[1], x, 1.0, sysm[2], sysm[3], orc.STATE_TRANSFER)
            assert abs(float(fg[0]) - Fo) < 1e-12 and float(fg[1:].abs().max()) == 0.0
