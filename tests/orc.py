"""ctypes loader for the CPU oracle (oracle/_build/libezpz_oracle.so) — test infrastructure only.

Builds the library on first use with oracle/Makefile.  Record constructors mirror the flat
64-byte layout documented in include/ezpz_b200.h.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libezpz_oracle.so")


class Rec(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("flags", C.c_uint32), ("ids", C.c_uint32 * 8),
                ("p0", C.c_double), ("p1", C.c_double), ("weight", C.c_double)]


class Cfg(C.Structure):
    _fields_ = [("max_iterations", C.c_uint64), ("residual_tolerance", C.c_double),
                ("step_tolerance", C.c_double), ("initial_lambda", C.c_double)]


class OrcOutcome(C.Structure):
    _fields_ = [("final_values", C.POINTER(C.c_double)), ("unsatisfied", C.POINTER(C.c_uint64)),
                ("n_unsatisfied", C.c_uint32), ("degen_count", C.POINTER(C.c_uint32)),
                ("underconstrained", C.POINTER(C.c_uint32)), ("n_underconstrained", C.c_uint32),
                ("iterations", C.c_uint64), ("converged", C.c_uint32), ("priority_solved", C.c_uint32),
                ("num_vars", C.c_uint32), ("num_eqs", C.c_uint32), ("err_constraint_id", C.c_uint64),
                ("err_variable", C.c_uint32)]


REC_DTYPE = np.dtype([("kind", "<u4"), ("flags", "<u4"), ("ids", "<u4", (8,)), ("p0", "<f8"),
                      ("p1", "<f8"), ("weight", "<f8")])
assert REC_DTYPE.itemsize == 64

_lib = None


def build():
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        src_newer = (not os.path.exists(LIB_PATH)) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(LIB_PATH)
            for f in ("ezpz_oracle.cpp", "constraints_ref.h", "libm_port.h"))
        if src_newer:
            build()
        L = C.CDLL(LIB_PATH)
        L.orc_fn_hypot.restype = C.c_double
        L.orc_fn_hypot.argtypes = [C.c_double, C.c_double]
        for f in ("orc_fn_sin", "orc_fn_cos"):
            getattr(L, f).restype = C.c_double
            getattr(L, f).argtypes = [C.c_double]
        L.orc_fn_atan2.restype = C.c_double
        L.orc_fn_atan2.argtypes = [C.c_double, C.c_double]
        L.orc_hardware_threads.restype = C.c_uint32
        _lib = L
    return _lib


def default_cfg(max_iterations=35, residual_tolerance=1e-8, step_tolerance=1e-12, initial_lambda=1e-9):
    return Cfg(max_iterations, residual_tolerance, step_tolerance, initial_lambda)


def as_recs(recs):
    """list of Rec / numpy structured array -> contiguous numpy structured array."""
    if isinstance(recs, np.ndarray):
        assert recs.dtype == REC_DTYPE
        return np.ascontiguousarray(recs)
    arr = np.zeros(len(recs), dtype=REC_DTYPE)
    for i, r in enumerate(recs):
        arr[i]["kind"] = r.kind
        arr[i]["flags"] = r.flags
        arr[i]["ids"] = list(r.ids)
        arr[i]["p0"] = r.p0
        arr[i]["p1"] = r.p1
        arr[i]["weight"] = r.weight
    return arr


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class Result:
    pass


def _mk_outcome(n_cons, n_vars):
    fv = np.zeros(max(n_vars, 1), dtype=np.float64)
    un = np.zeros(max(n_cons, 1), dtype=np.uint64)
    dg = np.zeros(max(n_cons, 1), dtype=np.uint32)
    uc = np.zeros(max(n_vars, 1), dtype=np.uint32)
    o = OrcOutcome()
    o.final_values = _p(fv, C.c_double)
    o.unsatisfied = _p(un, C.c_uint64)
    o.degen_count = _p(dg, C.c_uint32)
    o.underconstrained = _p(uc, C.c_uint32)
    return o, fv, un, dg, uc


def _result(rc, o, fv, un, dg, uc, n_cons, n_vars):
    r = Result()
    r.rc = rc
    r.final_values = fv[:n_vars].copy()
    r.unsatisfied = [int(v) for v in un[:o.n_unsatisfied]]
    r.degen_count = dg[:n_cons].copy()
    r.underconstrained = [int(v) for v in uc[:o.n_underconstrained]]
    r.iterations = int(o.iterations)
    r.converged = bool(o.converged)
    r.priority_solved = int(o.priority_solved)
    r.num_vars = int(o.num_vars)
    r.num_eqs = int(o.num_eqs)
    r.err_constraint_id = int(o.err_constraint_id)
    r.err_variable = int(o.err_variable)
    return r


def solve(recs, guesses, priorities=None, var_ids=None, cfg=None, analysis=False):
    """ezpz::solve / solve_analysis through the oracle (priority loop included)."""
    L = lib()
    recs = as_recs(recs)
    n_cons = len(recs)
    g = np.ascontiguousarray(guesses, dtype=np.float64)
    n_vars = len(g)
    cfg = cfg or default_cfg()
    pr = np.ascontiguousarray(priorities, dtype=np.uint32) if priorities is not None else None
    vi = np.ascontiguousarray(var_ids, dtype=np.uint32) if var_ids is not None else None
    o, fv, un, dg, uc = _mk_outcome(n_cons, n_vars)
    rc = L.orc_solve(recs.ctypes.data_as(C.c_void_p), _p(pr, C.c_uint32), C.c_uint32(n_cons),
                     _p(vi, C.c_uint32), _p(g, C.c_double), C.c_uint32(n_vars), C.byref(cfg),
                     C.c_int32(1 if analysis else 0), C.byref(o))
    return _result(rc, o, fv, un, dg, uc, n_cons, n_vars)


def solve_inner(recs, guesses, params=None, cfg=None, analysis=False, resolve_sides=True, trace=False):
    L = lib()
    recs = as_recs(recs)
    n_cons = len(recs)
    g = np.ascontiguousarray(guesses, dtype=np.float64)
    n_vars = len(g)
    cfg = cfg or default_cfg()
    po = np.ascontiguousarray(params, dtype=np.float64) if params is not None else None
    o, fv, un, dg, uc = _mk_outcome(n_cons, n_vars)
    cap = 4 * (int(cfg.max_iterations) + 1)
    tr = np.zeros(cap, dtype=np.float64)
    tl = C.c_uint32(0)
    rc = L.orc_solve_inner(recs.ctypes.data_as(C.c_void_p), C.c_uint32(n_cons), _p(g, C.c_double),
                           C.c_uint32(n_vars), _p(po, C.c_double), C.byref(cfg),
                           C.c_int32(1 if analysis else 0), C.c_int32(1 if resolve_sides else 0),
                           C.byref(o), _p(tr, C.c_double) if trace else None, C.c_uint32(cap), C.byref(tl))
    r = _result(rc, o, fv, un, dg, uc, n_cons, n_vars)
    r.trace = tr[:tl.value].reshape(-1, 4).copy() if trace else None
    return r


def solve_inner_ordered(recs, guesses, elim_order=None, sum_chunk=0, cfg=None, resolve_sides=True):
    """solve_inner with the two large-system knobs of the oracle: an elimination order for the Cholesky and the
    chunk length of the sum-of-squares fold (what Structure.ordering() reports for the CUDA large path)."""
    L = lib()
    recs = as_recs(recs)
    n_cons = len(recs)
    g = np.ascontiguousarray(guesses, dtype=np.float64)
    n_vars = len(g)
    cfg = cfg or default_cfg()
    eo = None if elim_order is None else np.ascontiguousarray(elim_order, dtype=np.uint32)
    o, fv, un, dg, uc = _mk_outcome(n_cons, n_vars)
    rc = L.orc_solve_inner_ordered(recs.ctypes.data_as(C.c_void_p), C.c_uint32(n_cons), _p(g, C.c_double),
                                   C.c_uint32(n_vars), C.byref(cfg), C.c_int32(1 if resolve_sides else 0),
                                   _p(eo, C.c_uint32), C.c_uint32(int(sum_chunk)), C.byref(o))
    return _result(rc, o, fv, un, dg, uc, n_cons, n_vars)


def pattern(recs, n_vars):
    L = lib()
    recs = as_recs(recs)
    n_cons = len(recs)
    m = C.c_uint32(0)
    nnz = C.c_uint64(0)
    rp = recs.ctypes.data_as(C.c_void_p)
    rc = L.orc_pattern(rp, C.c_uint32(n_cons), C.c_uint32(n_vars), C.byref(m), C.byref(nnz), None, None,
                       None, None, None)
    if rc != 0:
        return rc, None
    cp = np.zeros(n_vars + 1, np.uint32)
    ri = np.zeros(nnz.value, np.uint32)
    rpz = np.zeros(m.value + 1, np.uint32)
    ci = np.zeros(nnz.value, np.uint32)
    c0 = np.zeros(n_cons + 1, np.uint32)
    rc = L.orc_pattern(rp, C.c_uint32(n_cons), C.c_uint32(n_vars), C.byref(m), C.byref(nnz),
                       _p(cp, C.c_uint32), _p(ri, C.c_uint32), _p(rpz, C.c_uint32), _p(ci, C.c_uint32),
                       _p(c0, C.c_uint32))
    return rc, dict(m=m.value, nnz=nnz.value, csc_col_ptr=cp, csc_row_idx=ri, csr_row_ptr=rpz,
                    csr_col_idx=ci, cons_row0=c0)


def pattern_chol(recs, n_vars):
    L = lib()
    recs = as_recs(recs)
    na = C.c_uint64(0)
    nl = C.c_uint64(0)
    rp = recs.ctypes.data_as(C.c_void_p)
    rc = L.orc_pattern_chol(rp, C.c_uint32(len(recs)), C.c_uint32(n_vars), C.byref(na), C.byref(nl), None,
                            None, None, None)
    assert rc == 0
    arp = np.zeros(n_vars + 1, np.uint32)
    ac = np.zeros(na.value, np.uint32)
    lrp = np.zeros(n_vars + 1, np.uint32)
    lc = np.zeros(nl.value, np.uint32)
    rc = L.orc_pattern_chol(rp, C.c_uint32(len(recs)), C.c_uint32(n_vars), C.byref(na), C.byref(nl),
                            _p(arp, C.c_uint32), _p(ac, C.c_uint32), _p(lrp, C.c_uint32), _p(lc, C.c_uint32))
    assert rc == 0
    return dict(a_row_ptr=arp, a_col=ac, l_row_ptr=lrp, l_col=lc)


def evaluate(recs, n_vars, x):
    L = lib()
    recs = as_recs(recs)
    rc, pat = pattern(recs, n_vars)
    assert rc == 0, rc
    x = np.ascontiguousarray(x, dtype=np.float64)
    r = np.zeros(pat["m"], np.float64)
    j = np.zeros(pat["nnz"], np.float64)
    dg = np.zeros(len(recs), np.uint8)
    rc = L.orc_eval(recs.ctypes.data_as(C.c_void_p), C.c_uint32(len(recs)), C.c_uint32(n_vars),
                    _p(x, C.c_double), _p(r, C.c_double), _p(j, C.c_double), _p(dg, C.c_uint8))
    assert rc == 0
    return r, j, dg, pat


def solve_batch(recs, n_vars, guesses, params=None, cfg=None, nthreads=0, hoist=False, verdicts=False):
    """Batch of oracle solves on host threads -> (finals, iterations, status) [+ (unsat_mask, under_mask) with verdicts=True:
    per problem the unsatisfied constraint bits and the underconstrained variable bits of solve_analysis]."""
    L = lib()
    recs = as_recs(recs)
    g = np.ascontiguousarray(guesses, dtype=np.float64).reshape(-1, n_vars)
    B = g.shape[0]
    cfg = cfg or default_cfg()
    po = np.ascontiguousarray(params, dtype=np.float64) if params is not None else None
    fin = np.zeros_like(g)
    it = np.zeros(B, np.uint32)
    st = np.zeros(B, np.uint8)
    um = np.zeros((B, (len(recs) + 31) // 32), np.uint32) if verdicts else None
    vm = np.zeros((B, (n_vars + 31) // 32), np.uint32) if verdicts else None
    rc = L.orc_solve_batch_verdicts(recs.ctypes.data_as(C.c_void_p), C.c_uint32(len(recs)), C.c_uint32(n_vars),
                                    C.byref(cfg), C.c_uint64(B), _p(g, C.c_double), _p(po, C.c_double),
                                    _p(fin, C.c_double), _p(it, C.c_uint32), _p(st, C.c_uint8), C.c_uint32(nthreads),
                                    C.c_int32(1 if hoist else 0), _p(um, C.c_uint32), _p(vm, C.c_uint32))
    assert rc == 0, rc
    if verdicts:
        return fin, it, st, um, vm
    return fin, it, st


def sincos(rad):
    L = lib()
    return L.orc_fn_sin(rad), L.orc_fn_cos(rad)
