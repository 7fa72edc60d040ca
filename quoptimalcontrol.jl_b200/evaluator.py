"""GrapeEvaluator: one persistent device handle = the reference's work arrays (init_GRAPE,
/root/reference/src/grape_tools.jl:4-16) + the body of the Optim.only_fg! closure
(/root/reference/src/solve.jl:75-100, :164-196), executed by libqocgrape.so on the GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _colmajor(mats, D):
    """Stack of matrices -> contiguous buffer of column-major complex128 D x D blocks."""
    a = np.asarray(mats, dtype=np.complex128)
    if a.shape[-2:] != (D, D):
        raise ValueError(f"expected matrices of shape ({D},{D}), got {a.shape}")
    return np.ascontiguousarray(np.swapaxes(a, -1, -2))


class GrapeEvaluator:
    """members: list of (A, B, Xi, Xt) tuples (one per ensemble member; a plain Problem is one member).
    x has shape (K, N) like the reference's control_array, or (R, K, N) for a multi-start batch."""

    def __init__(self, members, T, n_slices, sys_type, wts=None, gradient="first_order",
                 convention="inplace", n_pulses=1, device=0, expm_theta=0.0, pure_state=True, devices=None,
                 penalty=(0.0, 0.0)):
        """devices: list of CUDA ordinals -> single-process multi-device handle (members block-sharded over them);
        penalty: (w_amp, w_var) weights of the C3 / C4 control penalties (src/cost_functions.jl:29-39)."""
        self._h = C.c_void_p()
        self._lib = _lib.load()
        M = len(members)
        A0, B0, Xi0, Xt0 = members[0]
        D = np.asarray(A0).shape[0]
        K = len(B0)
        self.D, self.K, self.N, self.M, self.R = D, K, int(n_slices), M, int(n_pulses)
        code = sys_type if isinstance(sys_type, int) else sys_type.code
        desc = _lib.QocDesc(sys_type=code, D=D, K=K, N=self.N, M=M, R=self.R, T=float(T),
                            gradient={"first_order": _lib.GRAD_FIRST_ORDER, "exact": _lib.GRAD_EXACT}[gradient],
                            convention={"inplace": _lib.REF_INPLACE, "static": _lib.REF_STATIC}[convention],
                            device=int(device), expm_theta=float(expm_theta),
                            flags=0 if pure_state else _lib.QOC_FLAG_NO_PURE_STATE)
        if devices is not None and len(devices) > 1:
            if len(devices) > _lib.MAX_DEVICES:
                raise ValueError(f"at most {_lib.MAX_DEVICES} devices")
            desc.n_devices = len(devices)
            for i, dv in enumerate(devices):
                desc.device_ids[i] = int(dv)
        elif devices is not None and len(devices) == 1:
            desc.device = int(devices[0])
        rc = self._lib.qoc_create(C.byref(self._h), C.byref(desc))
        if rc != _lib.QOC_OK:
            msg = self._lib.qoc_last_error(None).decode()
            self._h = C.c_void_p()
            raise _lib.QocError(rc, msg)
        A = _colmajor([m[0] for m in members], D)
        B = _colmajor([[b for b in m[1]] for m in members], D) if K else np.zeros(1, dtype=np.complex128)
        Xi = _colmajor([m[2] for m in members], D)
        Xt = _colmajor([m[3] for m in members], D)
        w = None if wts is None else np.ascontiguousarray(np.asarray(wts, dtype=np.float64))
        if w is not None and w.shape != (M,):
            raise ValueError("wts must have one weight per member")
        self._check(self._lib.qoc_set_system(self._h, A.ctypes.data, B.ctypes.data, Xi.ctypes.data, Xt.ctypes.data,
                                             None if w is None else w.ctypes.data, 0))
        if penalty[0] or penalty[1]:
            self.set_penalty(*penalty)

    # ------------------------------------------------------------------ helpers
    def _check(self, rc):
        if rc != _lib.QOC_OK:
            raise _lib.QocError(rc, self._lib.qoc_last_error(self._h).decode())

    def _pack_x(self, x):
        x = np.asarray(x, dtype=np.float64)
        if x.ndim == 2:
            x = x[None]
        if x.shape != (self.R, self.K, self.N):
            raise ValueError(f"pulse must have shape ({self.R},{self.K},{self.N}) or ({self.K},{self.N}), got {x.shape}")
        return np.ascontiguousarray(np.swapaxes(x, 1, 2))      # [R][N][K] == Julia K x N column-major

    # ------------------------------------------------------------------ API
    def eval(self, x, want_grad=True):
        """Returns (F, G): floats/arrays for R == 1 and 2-D x, else F[R], G[R, K, N]."""
        single = np.asarray(x).ndim == 2
        xb = self._pack_x(x)
        F = np.empty(self.R)
        G = np.empty((self.R, self.N, self.K)) if want_grad else None
        self._check(self._lib.qoc_eval(self._h, xb.ctypes.data, F.ctypes.data, None if G is None else G.ctypes.data))
        Gk = None if G is None else np.swapaxes(G, 1, 2)
        if single:
            return float(F[0]), (None if Gk is None else np.ascontiguousarray(Gk[0]))
        return F, Gk

    def set_penalty(self, w_amp=0.0, w_var=0.0):
        """F += w_amp*C3(x) + w_var*C4(x) (and the gradient) on every following evaluation."""
        self._check(self._lib.qoc_set_penalty(self._h, float(w_amp), float(w_var)))

    def eval_values(self, xs):
        """Fidelity-only evaluation of a batch of candidate pulses xs[R, K, N] in one call (gradient-free callers:
        a Nelder-Mead simplex, dCRAB candidates).  Returns F[R]."""
        F, _ = self.eval(xs, want_grad=False)
        return np.atleast_1d(F)

    def eval_allreduce(self, x, want_grad=True):
        """Like eval(), through qoc_eval_allreduce: every rank passes the same x and gets the sum over all ranks."""
        single = np.asarray(x).ndim == 2
        xb = self._pack_x(x)
        F = np.empty(self.R)
        G = np.empty((self.R, self.N, self.K)) if want_grad else None
        self._check(self._lib.qoc_eval_allreduce(self._h, xb.ctypes.data, F.ctypes.data, None if G is None else G.ctypes.data))
        Gk = None if G is None else np.swapaxes(G, 1, 2)
        if single:
            return float(F[0]), (None if Gk is None else np.ascontiguousarray(Gk[0]))
        return F, Gk

    def eval_raw(self, xb, F, G=None, allreduce=False):
        """The bare C-ABI call on caller-owned buffers in the ABI's own layout (what a Julia `ccall` passes): xb [R][N][K]
        float64 C-contiguous (= Julia's K x N column-major control_array per pulse), F [R], G [R][N][K] or None.  No
        per-call allocation or transposition on the Python side."""
        fn = self._lib.qoc_eval_allreduce if allreduce else self._lib.qoc_eval
        self._check(fn(self._h, xb.ctypes.data, F.ctypes.data, None if G is None else G.ctypes.data))

    def raw_caller(self, xb, F, G=None, allreduce=False):
        """eval_raw with the pointer marshalling done once: returns a zero-argument callable for tight loops (the per-call
        Python cost is then one ctypes foreign call, as close as Python gets to the Julia `ccall`)."""
        fn = self._lib.qoc_eval_allreduce if allreduce else self._lib.qoc_eval
        h, px, pF, pG = self._h, xb.ctypes.data, F.ctypes.data, None if G is None else G.ctypes.data
        keep = (xb, F, G)                      # the closure keeps the buffers alive

        def call():
            rc = fn(h, px, pF, pG)
            if rc != _lib.QOC_OK:
                self._check(rc)
            return keep
        return call

    def eval_device(self, x_dev_ptr, fg_dev_ptr, want_grad=True, stream=None):
        """Asynchronous evaluation on device pointers (ints): x [R][N][K], FG [R][1 + N*K]."""
        self._check(self._lib.qoc_eval_device(self._h, x_dev_ptr, fg_dev_ptr, int(want_grad), stream))

    # ---- multi-GPU: one-shot all-reduce over NVLink peer memory (one process per GPU) ----
    def comm_export(self) -> bytes:
        """Allocate the exchange buffer; returns its 64-byte CUDA IPC handle (to be all-gathered by the caller)."""
        buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        self._check(self._lib.qoc_comm_export(self._h, buf))
        return buf.raw

    def comm_connect(self, world, rank, handles):
        """handles: list of `world` IPC handles (bytes) in rank order, own entry included."""
        blob = b"".join(handles)
        if len(blob) != world * _lib.IPC_HANDLE_BYTES:
            raise ValueError("expected one 64-byte handle per rank")
        self._check(self._lib.qoc_comm_connect(self._h, int(world), int(rank), blob))

    def eval_allreduce_device(self, x_dev_ptr, fg_dev_ptr, want_grad=True, stream=None):
        """qoc_eval_device + fused one-shot all-reduce: FG receives the sum over all ranks (async on `stream`)."""
        self._check(self._lib.qoc_eval_allreduce_device(self._h, x_dev_ptr, fg_dev_ptr, int(want_grad), stream))

    # ---- slice-parallel evaluation of one large instance (one process per GPU, qoc_slice_*) ----
    def slice_export(self) -> bytes:
        buf = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        self._check(self._lib.qoc_slice_export(self._h, buf))
        return buf.raw

    def slice_connect(self, world, rank, handles, Xi, Xt):
        """handles: `world` IPC handles in rank order; Xi, Xt: the GLOBAL initial / target operators of the problem."""
        blob = b"".join(handles)
        if len(blob) != world * _lib.IPC_HANDLE_BYTES:
            raise ValueError("expected one 64-byte handle per rank")
        xi, xt = _colmajor([Xi], self.D), _colmajor([Xt], self.D)
        self._check(self._lib.qoc_slice_connect(self._h, int(world), int(rank), blob, xi.ctypes.data, xt.ctypes.data))

    def eval_slice(self, x_local, want_grad=True):
        """x_local[K, N_r]: this rank's slices.  Returns (F of the whole problem, G[K, N_r] of this rank's slices)."""
        xb = self._pack_x(x_local)
        F = np.empty(1)
        G = np.empty((self.N, self.K)) if want_grad else None
        self._check(self._lib.qoc_eval_slice(self._h, xb.ctypes.data, F.ctypes.data, None if G is None else G.ctypes.data))
        return float(F[0]), (None if G is None else np.ascontiguousarray(G.T))

    def minimize_lbfgs(self, x0, max_iters=0, history=0, g_tol=0.0, f_tol=-1.0, max_linesearch=0, linesearch="hagerzhang"):
        """L-BFGS inside the library (qoc_minimize_lbfgs): returns (x[K, N], result dict).  Single pulse only.
        linesearch: "hagerzhang" (Optim.LBFGS's default) or "backtracking"."""
        xb = self._pack_x(x0)
        out = np.empty_like(xb)
        opt = _lib.QocLbfgsOptions(max_iters=int(max_iters), history=int(history), g_tol=float(g_tol), f_tol=float(f_tol),
                                   max_linesearch=int(max_linesearch), linesearch={"hagerzhang": 0, "backtracking": 1}[linesearch])
        res = _lib.QocLbfgsResult()
        self._check(self._lib.qoc_minimize_lbfgs(self._h, xb.ctypes.data, C.byref(opt), out.ctypes.data, C.byref(res)))
        return np.ascontiguousarray(np.swapaxes(out, 1, 2)[0]), {f: getattr(res, f) for f, _ in res._fields_}

    def total_propagator(self, x):
        """pw_evolve with U0 = I (src/timeevolution.jl:28-39): U[R, M, D, D] (squeezed for R = M = 1)."""
        xb = self._pack_x(x)
        U = np.empty((self.R, self.M, self.D, self.D), dtype=np.complex128)
        self._check(self._lib.qoc_total_propagator(self._h, xb.ctypes.data, U.ctypes.data))
        U = np.swapaxes(U, -1, -2)
        return U[0, 0].copy() if (self.R == 1 and self.M == 1) else U

    def set_states(self, Xi, Xt):
        """Replace the initial / target operators of a single-member handle (slice-parallel use, qoc_set_states)."""
        xi = _colmajor([Xi], self.D)
        xt = _colmajor([Xt], self.D)
        self._check(self._lib.qoc_set_states(self._h, xi.ctypes.data, xt.ctypes.data, 0))

    def eval_continue(self, want_grad=True):
        """(F, G) for the pulse of the immediately preceding total_propagator() call, reusing its propagators."""
        F = np.empty(1)
        G = np.empty((self.N, self.K)) if want_grad else None
        self._check(self._lib.qoc_eval_continue(self._h, F.ctypes.data, None if G is None else G.ctypes.data))
        return float(F[0]), (None if G is None else np.ascontiguousarray(G.T))

    def propagators(self, x, what="propagator"):
        """pw_prop_save! / pw_ham_save! / pw_gen_save!: out[R, M, N, D, D] (squeezed for R = M = 1)."""
        mode = {"propagator": 0, "hamiltonian": 1, "generator": 2}[what]
        xb = self._pack_x(x)
        P = np.empty((self.R, self.M, self.N, self.D, self.D), dtype=np.complex128)
        self._check(self._lib.qoc_propagators(self._h, xb.ctypes.data, P.ctypes.data, mode))
        P = np.swapaxes(P, -1, -2)
        return P[0, 0].copy() if (self.R == 1 and self.M == 1) else P

    def stats(self):
        st = _lib.QocStats()
        self._check(self._lib.qoc_get_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in st._fields_}

    def close(self):
        if self._h:
            self._lib.qoc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
