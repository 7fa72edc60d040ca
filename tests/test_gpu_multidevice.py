"""Single-process multi-device handle (qoc_desc.n_devices / device_ids), the fused reduction + all-reduce kernel, the
control penalties and the batched fidelity-only entry.  Everything here also runs on a ONE-GPU box: device ordinals may
repeat (several shards on one device) and a one-rank communicator is a valid communicator; with >= 2 GPUs the same tests
additionally spread the shards over real devices."""
import os

import numpy as np
import pytest
import torch

import quoptimalcontrol_jl_b200 as qoc
from oracle import c_oracle, grape_oracle as orc
from conftest import assert_parity, random_system

pytestmark = pytest.mark.gpu


def _devices(n):
    have = max(1, torch.cuda.device_count())
    return [i % have for i in range(n)]


@pytest.mark.parametrize("n_dev", [2, 3, 8])
@pytest.mark.parametrize("D,M,sys_name", [(8, 37, "unitary"), (4, 9, "state"), (16, 5, "coherence")])
def test_multi_device_handle_matches_oracle(n_dev, D, M, sys_name, monkeypatch):
    st = {"state": orc.STATE_TRANSFER, "unitary": orc.UNITARY_GATE, "coherence": orc.COHERENCE_TRANSFER}[sys_name]
    K, N, T = 3, 24, 1.1
    members = [random_system(D, K, seed=600 + k, hermitian=(sys_name != "coherence"), unitary_targets=(sys_name == "unitary")) for k in range(M)]
    wts = np.random.default_rng(M).random(M)
    x = np.random.default_rng(D).uniform(-1, 1, (K, N))
    Fo, Go = orc.ensemble_fom_and_gradient(members, wts, x, T, st)
    for peer in ("1", "0"):                      # direct peer-memory loads / staged peer copies
        monkeypatch.setenv("QOC_MULTI_PEER", peer)
        with qoc.GrapeEvaluator(members, T, N, st, wts=wts, devices=_devices(n_dev)) as ev:
            for _ in range(3):                   # capture, then two replays of the multi-device graph
                F, G = ev.eval(x)
            assert_parity(F, G, Fo, Go)
            F0, _ = ev.eval(x, want_grad=False)
            assert_parity(F0, None, Fo, None)
            s = ev.stats()
            assert s["n_evals"] == 4 and s["launches_last_eval"] >= min(n_dev, M) + 1 and s["workspace_bytes"] > 0


def test_multi_device_more_devices_than_members_and_batches():
    """n_devices > M collapses to M shards; R > 1 pulses per call work through the same graph."""
    D, K, N, T, M, R = 8, 2, 12, 0.8, 3, 4
    members = [random_system(D, K, seed=700 + k) for k in range(M)]
    wts = np.array([0.2, 0.3, 0.5])
    xs = np.random.default_rng(5).uniform(-1, 1, (R, K, N))
    with qoc.GrapeEvaluator(members, T, N, orc.STATE_TRANSFER, wts=wts, n_pulses=R, devices=_devices(8)) as ev:
        F, G = ev.eval(xs)
    for r in range(R):
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, xs[r], T, orc.STATE_TRANSFER)
        assert_parity(F[r], G[r], Fo, Go)


def test_multi_device_cfg4_full_and_lbfgs():
    """The full 4096-member config through one multi-device handle (8 shards), then the native L-BFGS on a small ensemble."""
    cfg = qoc.configs.config4()
    Fo, Go = c_oracle.eval_ensemble(cfg["members"], cfg["wts"], cfg["x"], cfg["T"], cfg["sys_type"], 0, os.cpu_count() or 1)
    with qoc.GrapeEvaluator(cfg["members"], cfg["T"], cfg["N"], cfg["sys_type"], wts=cfg["wts"], devices=_devices(8)) as ev:
        F, G = ev.eval(cfg["x"])
    assert_parity(F, G, Fo, Go)
    small = qoc.configs.config4(N=30, grid=3)
    # exact gradient: the true derivative of the C1 functional, so a descent method must make progress from a random pulse
    with qoc.GrapeEvaluator(small["members"], small["T"], small["N"], small["sys_type"], wts=small["wts"], gradient="exact",
                            devices=_devices(2)) as ev:
        F0, _ = ev.eval(small["x"])
        x, info = ev.minimize_lbfgs(small["x"], max_iters=15)
        assert info["minimum"] < F0 and info["f_calls"] >= info["iterations"]


def test_multi_device_rejects_single_device_entries():
    members = [random_system(4, 2, seed=k) for k in range(4)]
    with qoc.GrapeEvaluator(members, 1.0, 8, orc.STATE_TRANSFER, devices=_devices(2)) as ev:
        with pytest.raises(qoc.QocError) as e:
            ev.total_propagator(np.zeros((2, 8)))
        assert e.value.status == qoc._lib.QOC_EUNSUPPORTED
        with pytest.raises(qoc.QocError):
            ev.comm_export()
    with pytest.raises(qoc.QocError):
        qoc.GrapeEvaluator(members, 1.0, 8, orc.STATE_TRANSFER, devices=[0, 99])


@pytest.mark.parametrize("D,M", [(8, 150), (8, 3), (4, 700), (32, 2)])
def test_one_rank_communicator_runs_the_fused_allreduce(D, M):
    """world = 1: the fused member-reduction + all-reduce kernel (device-side epoch, graph replay) must reproduce qoc_eval,
    over several epochs (both halves of the exchange buffer), through the host-buffer and the device-pointer entries."""
    K, N, T = 2, 16, 0.9
    members = [random_system(D, K, seed=800 + (k % 5)) for k in range(M)]
    wts = np.random.default_rng(M).random(M) / M
    with qoc.GrapeEvaluator(members, T, N, orc.STATE_TRANSFER, wts=wts, pure_state=False) as ev:
        ev.comm_connect(1, 0, [ev.comm_export()])
        with pytest.raises(qoc.QocError):
            ev.comm_connect(1, 0, [ev.comm_export()])         # once per handle
        x_dev = torch.zeros(N * K, dtype=torch.float64, device="cuda")
        fg = torch.zeros(N * K + 1, dtype=torch.float64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for it in range(5):
            x = np.random.default_rng(it).uniform(-1, 1, (K, N))
            Fr, Gr = ev.eval(x)
            Fa, Ga = ev.eval_allreduce(x)
            assert abs(Fa - Fr) <= 1e-13 * max(1, abs(Fr)) and np.max(np.abs(Ga - Gr)) <= 1e-13 * np.max(np.abs(Gr))
            x_dev.copy_(torch.from_numpy(np.ascontiguousarray(x.T).ravel()))
            ev.eval_allreduce_device(x_dev.data_ptr(), fg.data_ptr(), True, st)
            torch.cuda.synchronize()
            out = fg.cpu().numpy()
            assert abs(out[0] - Fr) <= 1e-13 * max(1, abs(Fr)) and np.max(np.abs(out[1:].reshape(N, K).T - Gr)) <= 1e-13 * np.max(np.abs(Gr))
            Fv, _ = ev.eval_allreduce(x, want_grad=False)
            assert abs(Fv - Fr) <= 1e-13 * max(1, abs(Fr))
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, x, T, orc.STATE_TRANSFER)
        assert_parity(Fa, Ga, Fo, Go)


def _penalties(x, wa, wv):
    """w_amp*C3 + w_var*C4 and the gradient (src/cost_functions.jl:29-39)."""
    d = np.diff(x, axis=1)
    F = wa * np.sum(x ** 2) + wv * np.sum(d ** 2)
    G = 2 * wa * x
    G[:, :-1] -= 2 * wv * d
    G[:, 1:] += 2 * wv * d
    return F, G


@pytest.mark.parametrize("D", [4, 32])
def test_control_penalties(D):
    K, N, T, M = 3, 10, 1.0, 2
    members = [random_system(D, K, seed=900 + k) for k in range(M)]
    wts = [0.4, 0.6]
    wa, wv = 0.03, 0.2
    xs = np.random.default_rng(2).uniform(-1, 1, (2, K, N))
    with qoc.GrapeEvaluator(members, T, N, orc.STATE_TRANSFER, wts=wts, n_pulses=2, penalty=(wa, wv), pure_state=False) as ev:
        F, G = ev.eval(xs)
        Fv = ev.eval_values(xs)
        for r in range(2):
            Fo, Go = orc.ensemble_fom_and_gradient(members, wts, xs[r], T, orc.STATE_TRANSFER)
            Fp, Gp = _penalties(xs[r], wa, wv)
            assert_parity(F[r], G[r], Fo + Fp, Go + Gp)
            assert abs(Fv[r] - (Fo + Fp)) < 1e-12
        ev.set_penalty(0.0, 0.0)                  # switching the penalties off drops the captured graphs
        F, G = ev.eval(xs)
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, xs[0], T, orc.STATE_TRANSFER)
        assert_parity(F[0], G[0], Fo, Go)
    with qoc.GrapeEvaluator(members, T, N, orc.STATE_TRANSFER, wts=wts, penalty=(wa, wv), devices=_devices(2)) as ev:
        F, G = ev.eval(xs[1])
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, xs[1], T, orc.STATE_TRANSFER)
        Fp, Gp = _penalties(xs[1], wa, wv)
        assert_parity(F, G, Fo + Fp, Go + Gp)      # applied once, after the device sum


def test_penalty_gradient_is_the_derivative():
    """Finite differences of the C3 / C4 terms (host check of the formulas the kernel implements)."""
    x = np.random.default_rng(0).uniform(-1, 1, (2, 7))
    F0, G = _penalties(x, 0.3, 0.7)
    for (c, t) in [(0, 0), (1, 3), (0, 6)]:
        e = np.zeros_like(x); e[c, t] = 1e-6
        fd = (_penalties(x + e, 0.3, 0.7)[0] - _penalties(x - e, 0.3, 0.7)[0]) / 2e-6
        assert abs(fd - G[c, t]) < 1e-8


@pytest.mark.parametrize("sys_name", ["state", "unitary", "coherence"])
@pytest.mark.parametrize("gradient", ["first_order", "exact"])
def test_native_slice_parallel_one_rank(sys_name, gradient):
    """world = 1: qoc_eval_slice (range propagator into the exchange buffer, flag round trip, continue with the global
    states) must equal the plain evaluation."""
    st = {"state": orc.STATE_TRANSFER, "unitary": orc.UNITARY_GATE, "coherence": orc.COHERENCE_TRANSFER}[sys_name]
    D, K, N, T = 32, 2, 9, 0.8
    A, B, Xi, Xt = random_system(D, K, seed=1000 + st, hermitian=(sys_name != "coherence"), unitary_targets=(sys_name == "unitary"))
    x = np.random.default_rng(3).uniform(-1, 1, (K, N))
    ev = qoc.NativeSliceParallelEvaluator(A, B, Xi, Xt, T, N, st, gradient=gradient)
    for _ in range(3):                       # both halves of the exchange buffer
        F, G = ev.eval(x)
    ev.close()
    if gradient == "exact":
        Fo, Go = orc.exact_fom_and_gradient(A, B, x, T, Xi, Xt, st)
    else:
        Fo, Go = orc.fom_and_gradient_grape(A, B, x, T, Xi, Xt, st)
    assert_parity(F, G, Fo, Go)


def test_solve_ensemble_on_a_multi_device_algorithm():
    """solve(ens, GPUGRAPE(devices = [...])): the dispatch surface reaches the multi-device handle (mirror of
    julia/GPUGRAPE.jl's `devices` keyword) and finds the same minimum as the single-device solve."""
    Sx = np.array([[0, 1], [1, 0]], dtype=complex) / 2
    Sy = np.array([[0, -1j], [1j, 0]], dtype=complex) / 2
    Sz = np.diag([0.5, -0.5]).astype(complex)
    rho0, rho1 = np.diag([1, 0]).astype(complex), np.diag([0, 1]).astype(complex)
    guess = np.random.default_rng(4).random((2, 20))
    p0 = qoc.Problem(B=[Sx, Sy], A=Sz, Xi=rho0, Xt=rho1, T=4.0, n_controls=2, guess=guess, sys_type=qoc.StateTransfer())
    ens = qoc.EnsembleProblem(prob=p0, n_ens=5, A_g=lambda k: Sz * (1 + 0.02 * (k - 2)), B_g=lambda k: [Sx, Sy],
                              XiG=lambda k: rho0, XtG=lambda k: rho1, wts=np.ones(5) / 5)
    a = qoc.solve(ens, qoc.GPUGRAPE(n_slices=20, devices=_devices(3)))
    b = qoc.solve(ens, qoc.GPUGRAPE(n_slices=20))
    # the shard sums differ from the single-device sum in the last bits, so the two optimisation paths drift apart: compare
    # the minima they reach, not the trajectories
    assert a.fidelity < 0.7501 and b.fidelity < 0.7501 and abs(a.fidelity - b.fidelity) < 1e-4


def test_one_rank_allreduce_with_several_pulses():
    """R > 1 through the fused reduction + all-reduce kernel (rows [R][chunks][NK+1]) and through a multi-device handle."""
    D, K, N, T, M, R = 8, 3, 70, 1.0, 160, 3
    members = [random_system(D, K, seed=1200 + (k % 7), unitary_targets=True) for k in range(M)]
    wts = np.random.default_rng(1).random(M) / M
    xs = np.random.default_rng(2).uniform(-1, 1, (R, K, N))
    with qoc.GrapeEvaluator(members, T, N, orc.UNITARY_GATE, wts=wts, n_pulses=R) as ev:
        ev.comm_connect(1, 0, [ev.comm_export()])
        for _ in range(2):
            F, G = ev.eval_allreduce(xs)
        Fr, Gr = ev.eval(xs)
    with qoc.GrapeEvaluator(members, T, N, orc.UNITARY_GATE, wts=wts, n_pulses=R, devices=_devices(3)) as ev:
        Fm, Gm = ev.eval(xs)
    for r in range(R):
        Fo, Go = orc.ensemble_fom_and_gradient(members, wts, xs[r], T, orc.UNITARY_GATE)
        assert_parity(F[r], G[r], Fo, Go)
        assert_parity(Fr[r], Gr[r], Fo, Go)
        assert_parity(Fm[r], Gm[r], Fo, Go)
