"""ctypes wrapper of the C restatement (oracle/grape_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _has_avx2():
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        return " avx2 " in txt and " fma " in txt
    except OSError:
        return False


def load():
    global _lib
    if _lib is None:
        name = "libqoc_oracle_avx2.so" if _has_avx2() else "libqoc_oracle.so"
        path = os.path.join(HERE, name)
        if not os.path.exists(path):
            from . import build as _b
            _b.build()
        lib = C.CDLL(path)
        lib.qoc_oracle_eval.restype = C.c_int
        lib.qoc_oracle_eval.argtypes = [C.c_int] * 5 + [C.c_double] + [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        lib.qoc_oracle_expm.restype = C.c_int
        lib.qoc_oracle_expm.argtypes = [C.c_int, C.c_void_p]
        lib.qoc_oracle_max_threads.restype = C.c_int
        _lib = lib
    return _lib


def _cm(mats):
    return np.ascontiguousarray(np.swapaxes(np.asarray(mats, dtype=np.complex128), -1, -2))


def eval_ensemble(members, wts, x, T, sys_type, variant=0, nthreads=1, want_grad=True):
    """Same contract as grape_oracle.ensemble_fom_and_gradient; x has shape (K, N)."""
    lib = load()
    M = len(members)
    D = np.asarray(members[0][0]).shape[0]
    K, N = x.shape
    A = _cm([m[0] for m in members]); B = _cm([list(m[1]) for m in members])
    Xi = _cm([m[2] for m in members]); Xt = _cm([m[3] for m in members])
    w = np.ascontiguousarray(np.asarray(wts if wts is not None else np.ones(M), dtype=np.float64))
    xb = np.ascontiguousarray(np.asarray(x, dtype=np.float64).T)
    F = C.c_double(0.0)
    G = np.zeros((N, K))
    rc = lib.qoc_oracle_eval(int(sys_type), D, K, N, M, float(T), A.ctypes.data, B.ctypes.data, Xi.ctypes.data,
                             Xt.ctypes.data, w.ctypes.data, xb.ctypes.data, int(variant), int(nthreads),
                             C.addressof(F), G.ctypes.data if want_grad else None)
    if rc != 0:
        raise MemoryError("qoc_oracle_eval: allocation failed")
    return F.value, (np.ascontiguousarray(G.T) if want_grad else None)


def expm(X):
    lib = load()
    D = X.shape[0]
    buf = _cm(X).copy()
    prods = lib.qoc_oracle_expm(D, buf.ctypes.data)
    return np.swapaxes(buf, -1, -2).copy(), prods


def max_threads():
    return load().qoc_oracle_max_threads()
