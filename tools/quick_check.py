"""Two-second, torch-free GPU check (numpy + ctypes only): a 160-member Pauli-type ensemble (D = 8, K = 6, 67 slices) through the
chunk-parallel closed-system kernels with every structure-dependent form active, against the C restatement.
`python tools/quick_check.py` from the repo root; writes gpurun_out/quick_check.txt.  Last run (end of round 2, after the host-side
structure analysis became one function): F err 3.2e-16, gradient rel err 2.7e-15, 14 launches, 1.4 s."""
import sys, time
t0 = time.time()
sys.path.insert(0, '.')
import numpy as np
import quoptimalcontrol_jl_b200 as qoc
from oracle import c_oracle
SX = np.array([[0, 1], [1, 0]], dtype=complex); SY = np.array([[0, -1j], [1j, 0]]); SZ = np.diag([1.0 + 0j, -1]); I2 = np.eye(2, dtype=complex)
def on(op, q):
    out = np.array([[1.0 + 0j]])
    for i in range(3): out = np.kron(out, op if i == q else I2)
    return out
rng = np.random.default_rng(19)
A0 = 0.5 * (on(SZ, 0) @ on(SZ, 1) + on(SZ, 1) @ on(SZ, 2))
B0 = [on(SX, 0) / 2, on(SY, 0) / 2, on(SX, 1) / 2, on(SY, 1) / 2, on(SX, 2) / 2, on(SY, 2) / 2]
Xt = np.linalg.qr(rng.standard_normal((8, 8)) + 1j * rng.standard_normal((8, 8)))[0]
M, N, T = 160, 67, 1.7
mem = [(A0 + 0.2 * rng.standard_normal() * np.diag(np.arange(8) - 4.0), [(1 + 0.05 * rng.standard_normal()) * b for b in B0], np.eye(8, dtype=complex), Xt) for _ in range(M)]
w = np.full(M, 1.0 / M); x = rng.uniform(-1, 1, (6, N))
Fo, Go = c_oracle.eval_ensemble(mem, w, x, T, 1, 0, 8)
with qoc.GrapeEvaluator(mem, T, N, 1, wts=w) as ev:
    F, G = ev.eval(x)
    st = ev.stats()
msg = "quick_check F err %.2e  G rel err %.2e  launches %d  %.1f s" % (abs(F - Fo), np.max(np.abs(G - Go)) / np.max(np.abs(Go)), st["launches_last_eval"], time.time() - t0)
print(msg); open("gpurun_out/quick_check.txt", "w").write(msg + "\n")
