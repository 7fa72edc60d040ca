"""ctypes binding of libqocgrape.so (include/qocgrape.h).  Loading fails loudly: there is no CPU fallback."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QOCGRAPE_LIB") or os.path.join(HERE, "libqocgrape.so")   # env override: tuning builds

QOC_OK, QOC_EINVAL, QOC_ECUDA, QOC_ENOMEM, QOC_EUNSUPPORTED = 0, 1, 2, 3, 4
QOC_FLAG_NO_PURE_STATE = 1
STATE_TRANSFER, UNITARY_GATE, COHERENCE_TRANSFER = 0, 1, 2
GRAD_FIRST_ORDER, GRAD_EXACT = 0, 1
REF_INPLACE, REF_STATIC = 0, 1
SHARED_A, SHARED_B, SHARED_XI, SHARED_XT = 1, 2, 4, 8
IPC_HANDLE_BYTES = 64
MAX_DEVICES = 16

EXPORTS = ["qoc_version", "qoc_create", "qoc_destroy", "qoc_set_system", "qoc_eval", "qoc_eval_device",
           "qoc_total_propagator", "qoc_propagators", "qoc_get_stats", "qoc_last_error",
           "qoc_comm_export", "qoc_comm_connect", "qoc_eval_allreduce_device", "qoc_minimize_lbfgs",
           "qoc_set_states", "qoc_eval_continue", "qoc_eval_allreduce", "qoc_set_penalty",
           "qoc_slice_export", "qoc_slice_connect", "qoc_eval_slice", "qoc_analyze_structure"]


class QocDesc(C.Structure):
    _fields_ = [("sys_type", C.c_int), ("D", C.c_int), ("K", C.c_int), ("N", C.c_int), ("M", C.c_int),
                ("R", C.c_int), ("T", C.c_double), ("gradient", C.c_int), ("convention", C.c_int),
                ("device", C.c_int), ("expm_theta", C.c_double), ("flags", C.c_int),
                ("n_devices", C.c_int), ("device_ids", C.c_int * MAX_DEVICES)]


class QocStats(C.Structure):
    _fields_ = [("n_evals", C.c_longlong), ("n_launches", C.c_longlong), ("launches_last_eval", C.c_int),
                ("gpu_ms_last_eval", C.c_float), ("workspace_bytes", C.c_longlong), ("path", C.c_int),
                ("main_kernel_ms_avg", C.c_float), ("main_kernel_samples", C.c_int)]


class QocLbfgsOptions(C.Structure):
    _fields_ = [("max_iters", C.c_int), ("history", C.c_int), ("g_tol", C.c_double), ("f_tol", C.c_double),
                ("max_linesearch", C.c_int), ("linesearch", C.c_int)]


class QocLbfgsResult(C.Structure):
    _fields_ = [("minimum", C.c_double), ("g_norm", C.c_double), ("iterations", C.c_int), ("f_calls", C.c_int),
                ("converged", C.c_int)]


class QocError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"libqocgrape status {status}: {message}")
        self.status = status


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python quoptimalcontrol.jl_b200/build.py` "
                          "(the CUDA library is the only implementation; there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    dp, vp = C.POINTER(C.c_double), C.c_void_p
    lib.qoc_version.restype = C.c_char_p
    lib.qoc_last_error.restype = C.c_char_p
    lib.qoc_last_error.argtypes = [vp]
    lib.qoc_create.argtypes = [C.POINTER(vp), C.POINTER(QocDesc)]
    lib.qoc_destroy.argtypes = [vp]
    lib.qoc_set_system.argtypes = [vp, vp, vp, vp, vp, vp, C.c_int]
    lib.qoc_eval.argtypes = [vp, vp, vp, vp]
    lib.qoc_eval_device.argtypes = [vp, vp, vp, C.c_int, vp]
    lib.qoc_total_propagator.argtypes = [vp, vp, vp]
    lib.qoc_propagators.argtypes = [vp, vp, vp, C.c_int]
    lib.qoc_get_stats.argtypes = [vp, C.POINTER(QocStats)]
    lib.qoc_minimize_lbfgs.argtypes = [vp, vp, C.POINTER(QocLbfgsOptions), vp, C.POINTER(QocLbfgsResult)]
    lib.qoc_comm_export.argtypes = [vp, vp]
    lib.qoc_comm_connect.argtypes = [vp, C.c_int, C.c_int, vp]
    lib.qoc_eval_allreduce_device.argtypes = [vp, vp, vp, C.c_int, vp]
    lib.qoc_eval_allreduce.argtypes = [vp, vp, vp, vp]
    lib.qoc_set_penalty.argtypes = [vp, C.c_double, C.c_double]
    lib.qoc_slice_export.argtypes = [vp, vp]
    lib.qoc_slice_connect.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    lib.qoc_eval_slice.argtypes = [vp, vp, vp, vp]
    lib.qoc_set_states.argtypes = [vp, vp, vp, C.c_int]
    lib.qoc_eval_continue.argtypes = [vp, vp, vp]
    for name in EXPORTS:
        if name not in ("qoc_version", "qoc_last_error"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib
