"""Small evaluations covering every kernel family, for compute-sanitizer (tools/sanitize.sh).  Checks parity too, so a
sanitizer-clean run is also a correct run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import quoptimalcontrol_jl_b200 as qoc  # noqa: E402
from oracle import grape_oracle as orc  # noqa: E402
from conftest import assert_parity, random_system  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"


def run(D, K, N, M, st, gradient="first_order", env=None, **kw):
    for k, v in (env or {}).items():
        os.environ[k] = v
    herm = st != orc.COHERENCE_TRANSFER
    members = [random_system(D, K, seed=10 * D + k, hermitian=herm, unitary_targets=(st == orc.UNITARY_GATE)) for k in range(M)]
    wts = np.linspace(0.5, 1.0, M) / M
    x = np.random.default_rng(D).uniform(-1, 1, (K, N))
    with qoc.GrapeEvaluator(members, 1.0, N, st, wts=wts, gradient=gradient, pure_state=False, **kw) as ev:
        if kw.get("devices") is None and gradient == "first_order":
            ev.comm_connect(1, 0, [ev.comm_export()])
            F, G = ev.eval_allreduce(x)
        else:
            F, G = ev.eval(x)
        F2, G2 = ev.eval(x)
    Fo, Go = (orc.ensemble_exact if gradient == "exact" else orc.ensemble_fom_and_gradient)(members, wts, x, 1.0, st)
    assert_parity(F, G, Fo, Go)
    assert_parity(F2, G2, Fo, Go)
    for k in (env or {}):
        os.environ.pop(k, None)
    print("ok", D, K, N, M, st, gradient, env or "", kw or "", flush=True)


if which in ("all", "small"):
    for D in (2, 4, 8, 16):
        run(D, 3, 12, 3, orc.UNITARY_GATE)
        run(D, 2, 12, 2, orc.STATE_TRANSFER, "exact")
    run(16, 2, 12, 1, orc.COHERENCE_TRANSFER)
    run(8, 3, 40, 3, orc.UNITARY_GATE, env={"QOC_CHUNKED": "1", "QOC_CHUNKS": "3"})      # chunk-parallel closed-system kernels
    run(8, 3, 40, 2, orc.COHERENCE_TRANSFER, env={"QOC_CHUNKED": "1", "QOC_CHUNKS": "3"})  # chunk-parallel general kernels
    run(4, 2, 40, 2, orc.STATE_TRANSFER, "exact", env={"QOC_PHASED": "1", "QOC_CHUNKS": "4"})  # slice-parallel pipeline
if which in ("all", "big"):
    run(64, 2, 6, 1, orc.STATE_TRANSFER)
    run(32, 2, 6, 2, orc.COHERENCE_TRANSFER)
if which in ("all", "multi"):
    run(8, 3, 12, 5, orc.UNITARY_GATE, devices=[0, 0])
    import torch
    if torch.cuda.device_count() >= 2:
        run(8, 3, 12, 5, orc.UNITARY_GATE, devices=[0, 1])
print("sanitize_cases done")
