"""ctypes binding of libezpz_b200.so (include/ezpz_b200.h).  There is no fallback: if the library is
missing, import fails with instructions to build it."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EZPZ_B200_LIB") or os.path.join(_HERE, "_lib", "libezpz_b200.so")  # (EZPZ_B200_LIB: A/B timing of two builds)


class Constraint(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("flags", C.c_uint32), ("ids", C.c_uint32 * 8),
                ("p0", C.c_double), ("p1", C.c_double), ("weight", C.c_double)]


class Config(C.Structure):
    _fields_ = [("max_iterations", C.c_uint64), ("residual_tolerance", C.c_double),
                ("step_tolerance", C.c_double), ("initial_lambda", C.c_double)]


class ErrorDetail(C.Structure):
    _fields_ = [("constraint_id", C.c_uint64), ("variable", C.c_uint32), ("reserved", C.c_uint32),
                ("a", C.c_uint64), ("b", C.c_uint64), ("message", C.c_char * 192)]


class BatchIO(C.Structure):
    _fields_ = [("guesses", C.c_void_p), ("params", C.c_void_p), ("final_values", C.c_void_p),
                ("iterations", C.c_void_p), ("status", C.c_void_p), ("unsat_mask", C.c_void_p),
                ("degen_count", C.c_void_p), ("jacobian", C.c_void_p), ("under_mask", C.c_void_p)]


class BatchJob(C.Structure):
    _fields_ = [("structure", C.c_void_p), ("batch", C.c_uint64), ("io", BatchIO), ("status", C.c_int32), ("reserved", C.c_int32)]


class OneIO(C.Structure):
    _fields_ = [("guesses", C.c_void_p), ("final_values", C.c_void_p), ("iterations", C.c_void_p),
                ("status", C.c_void_p), ("unsat_mask", C.c_void_p), ("degen_count", C.c_void_p),
                ("jacobian", C.c_void_p), ("path_used", C.c_void_p), ("lin_iters", C.c_void_p)]


class WarningRec(C.Structure):
    _fields_ = [("about_constraint", C.c_int64), ("kind", C.c_uint32), ("count", C.c_uint32),
                ("angle_deg", C.c_double)]


class OutcomeRec(C.Structure):
    _fields_ = [("final_values", C.c_void_p), ("unsatisfied", C.c_void_p), ("underconstrained", C.c_void_p),
                ("warnings", C.c_void_p), ("warnings_cap", C.c_uint32), ("n_warnings", C.c_uint32),
                ("n_unsatisfied", C.c_uint32), ("n_underconstrained", C.c_uint32), ("iterations", C.c_uint64),
                ("converged", C.c_uint32), ("priority_solved", C.c_uint32), ("num_vars", C.c_uint32),
                ("num_eqs", C.c_uint32), ("path_used", C.c_int32), ("reserved", C.c_uint32)]


REC_DTYPE = np.dtype([("kind", "<u4"), ("flags", "<u4"), ("ids", "<u4", (8,)), ("p0", "<f8"),
                      ("p1", "<f8"), ("weight", "<f8")])
assert REC_DTYPE.itemsize == 64 and C.sizeof(Constraint) == 64

# Every symbol include/ezpz_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
_SYMBOLS = [
    ("ezpz_b200_structure_create", C.c_int32, [_P, C.c_uint32, _P, C.c_uint32, C.POINTER(_P), C.POINTER(ErrorDetail)]),
    ("ezpz_b200_structure_destroy", None, [_P]),
    ("ezpz_b200_structure_extend", C.c_int32, [_P, _P, C.c_uint32, C.POINTER(_P), C.POINTER(ErrorDetail)]),
    ("ezpz_b200_structure_dims", C.c_int32, [_P, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64),
                                             C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    ("ezpz_b200_structure_pattern", C.c_int32, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    ("ezpz_b200_structure_pattern_a", C.c_int32, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(_P)]),
    ("ezpz_b200_structure_rows", C.c_int32, [_P, C.POINTER(_P)]),
    ("ezpz_b200_structure_fingerprint", C.c_uint64, [_P]),
    ("ezpz_b200_structure_batch_shape", C.c_int32, [_P, C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(C.c_uint32),
                                                    C.POINTER(C.c_uint32)]),
    ("ezpz_b200_structure_role_program", C.c_int32, [_P, C.c_uint32, C.c_uint32, _P, C.c_uint64, C.POINTER(C.c_uint64),
                                                     C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    ("ezpz_b200_structure_ordering", C.c_int32, [_P, C.POINTER(C.c_int32), C.POINTER(_P), C.POINTER(C.c_int32),
                                                 C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    ("ezpz_b200_context_create", C.c_int32, [C.c_int32, C.POINTER(_P), C.POINTER(ErrorDetail)]),
    ("ezpz_b200_context_destroy", None, [_P]),
    ("ezpz_b200_context_launches", C.c_uint64, [_P]),
    ("ezpz_b200_context_synchronize", C.c_int32, [_P]),
    ("ezpz_b200_solve_batch", C.c_int32, [_P, _P, C.POINTER(Config), C.c_uint64, C.POINTER(BatchIO), C.POINTER(ErrorDetail)]),
    ("ezpz_b200_solve_batch_device", C.c_int32, [_P, _P, C.POINTER(Config), C.c_uint64, C.POINTER(BatchIO), _P, C.POINTER(ErrorDetail)]),
    ("ezpz_b200_multi_create", C.c_int32, [_P, C.c_int32, C.POINTER(_P), C.POINTER(ErrorDetail)]),
    ("ezpz_b200_multi_destroy", None, [_P]),
    ("ezpz_b200_multi_device_count", C.c_int32, [_P]),
    ("ezpz_b200_multi_context", _P, [_P, C.c_int32]),
    ("ezpz_b200_multi_launches", C.c_uint64, [_P]),
    ("ezpz_b200_solve_batch_multi", C.c_int32, [_P, _P, C.POINTER(Config), C.c_uint64, C.POINTER(BatchIO), C.POINTER(ErrorDetail)]),
    ("ezpz_b200_solve_jobs_multi", C.c_int32, [_P, C.POINTER(Config), C.POINTER(BatchJob), C.c_uint32, C.POINTER(ErrorDetail)]),
    ("ezpz_b200_host_register", C.c_int32, [_P, C.c_uint64]),
    ("ezpz_b200_host_unregister", C.c_int32, [_P]),
    ("ezpz_b200_host_alloc", C.c_int32, [C.c_uint64, C.POINTER(_P)]),
    ("ezpz_b200_host_free", None, [_P]),
    ("ezpz_b200_shard_range", None, [C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    ("ezpz_b200_solve_one", C.c_int32, [_P, _P, C.POINTER(Config), C.POINTER(OneIO), C.POINTER(ErrorDetail)]),
    ("ezpz_b200_large_bench", C.c_int32, [_P, _P, _P, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                          C.POINTER(ErrorDetail)]),
    ("ezpz_b200_eval", C.c_int32, [_P, _P, _P, _P, _P, _P, _P, C.POINTER(ErrorDetail)]),
    ("ezpz_b200_freedom_analysis", C.c_int32, [_P, _P, C.c_uint64, _P, _P, C.POINTER(ErrorDetail)]),
    ("ezpz_b200_context_clear_cache", None, [_P]),
    ("ezpz_b200_freedom_analysis_device", C.c_int32, [_P, _P, C.c_uint64, _P, _P, _P, C.POINTER(ErrorDetail)]),
    ("ezpz_b200_solve", C.c_int32, [_P, _P, _P, _P, C.c_uint32, _P, _P, C.c_uint32, C.POINTER(Config), C.c_int32,
                                    C.POINTER(OutcomeRec), C.POINTER(ErrorDetail)]),
    ("ezpz_b200_solve_batch_priorities", C.c_int32, [_P, _P, _P, C.c_uint32, C.c_uint32, C.POINTER(Config), C.c_uint64, _P, _P, _P, _P,
                                                     _P, _P, _P, C.POINTER(ErrorDetail)]),
    ("ezpz_b200_angle_sincos", None, [C.c_double, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    ("ezpz_b200_hypot", C.c_double, [C.c_double, C.c_double]),
    ("ezpz_b200_config_default", None, [C.POINTER(Config)]),
    ("ezpz_b200_abi_version", C.c_uint32, []),
    ("ezpz_b200_status_name", C.c_char_p, [C.c_int32]),
    ("ezpz_b200_problem_parse", C.c_int32, [C.c_char_p, C.c_uint64, C.POINTER(_P), C.POINTER(ErrorDetail)]),
    ("ezpz_b200_problem_destroy", None, [_P]),
    ("ezpz_b200_problem_system", C.c_int32, [_P, C.POINTER(_P), C.POINTER(C.c_uint32), C.POINTER(_P),
                                             C.POINTER(C.c_uint32), C.POINTER(ErrorDetail)]),
    ("ezpz_b200_problem_count", C.c_uint32, [_P, C.c_int32]),
    ("ezpz_b200_problem_label", C.c_char_p, [_P, C.c_int32, C.c_uint32]),
    ("ezpz_b200_problem_angles_deg", C.c_int32, [_P, C.POINTER(_P)]),
]
SYMBOL_NAMES = [s[0] for s in _SYMBOLS]

_lib = None


def lib():
    """Load the library (once).  Raises with build instructions when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "ezpz_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, restype, argtypes in _SYMBOLS:
            fn = getattr(L, name)  # AttributeError here = the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = L
    return _lib


def ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def status_name(rc):
    return lib().ezpz_b200_status_name(rc).decode()
