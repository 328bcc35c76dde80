"""The reference's fuzz contract (fuzz/fuzz_targets/fuzz_target_1.rs:7-23): `solve` on ARBITRARY constraint requests and
guesses never panics — it returns an outcome or an error.  Here: arbitrary records (every kind, ids inside and far outside the
guesses, NaN / infinite / denormal / huge parameters and guesses, arbitrary priorities and weights) through the oracle
(`-m "not gpu"`) and through the C ABI on the GPU (`-m gpu`); every call must come back with a status, and wherever both sides
succeed on finite input the iteration count, the convergence flag, the unsatisfied list and the coordinates agree."""
import ctypes as C

import numpy as np
import pytest

import orc
import textual_twin

REC = textual_twin.REC_DTYPE
N_IDS = [7, 6, 4, 5, 4, 4, 4, 4, 8, 1, 2, 4, 3, 8, 6, 6, 6, 6, 6, 6, 8, 8, 6, 6, 6]


def arbitrary_f64(rng, wild):
    if not wild:
        return float(rng.uniform(-10, 10))
    k = rng.integers(8)
    if k == 0:
        return float(np.frombuffer(rng.bytes(8), dtype=np.float64)[0])  # any bit pattern
    if k == 1:
        return float(rng.choice([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, 1e308, -1e308, 1e-300]))
    return float(rng.uniform(-100, 100))


def arbitrary_setup(rng, wild):
    n_vars = int(rng.integers(0, 17))
    n_cons = int(rng.integers(0, 13))
    recs = np.zeros(n_cons, dtype=REC)
    for c in range(n_cons):
        kind = int(rng.integers(25))
        recs[c]["kind"] = kind
        recs[c]["flags"] = int(rng.integers(0, 3))
        ids = rng.integers(0, max(1, n_vars), 8)
        if wild and rng.integers(4) == 0:
            ids[rng.integers(8)] = rng.integers(0, 2 ** 32 - 1)  # far outside the guesses
        recs[c]["ids"][:N_IDS[kind]] = ids[:N_IDS[kind]]
        recs[c]["p0"] = arbitrary_f64(rng, wild)
        recs[c]["p1"] = arbitrary_f64(rng, wild)
        recs[c]["weight"] = 1.0 if rng.integers(3) else abs(arbitrary_f64(rng, wild))
    prios = rng.integers(0, 3, n_cons).astype(np.uint32)
    guesses = np.array([arbitrary_f64(rng, wild) for _ in range(n_vars)], dtype=np.float64)
    return recs, prios, guesses


def same(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def gpu_solve(recs, prios, guesses):
    import ezpz_b200 as ez
    from ezpz_b200 import native
    n_cons, n_vars = len(recs), len(guesses)
    fv, un = np.zeros(max(1, n_vars)), np.zeros(max(1, n_cons), np.uint64)
    warr = (native.WarningRec * (2 * n_cons + 8))()
    out = native.OutcomeRec()
    out.final_values, out.unsatisfied, out.underconstrained = fv.ctypes.data, un.ctypes.data, None
    out.warnings, out.warnings_cap = C.addressof(warr), 2 * n_cons + 8
    det = native.ErrorDetail()
    cfg = ez.Config()._native()
    rc = native.lib().ezpz_b200_solve(ez.default_context().handle, native.ptr(recs) if n_cons else None, native.ptr(prios) if n_cons else None,
                                      None, n_cons, None, native.ptr(guesses) if n_vars else None, n_vars, C.byref(cfg), 0,
                                      C.byref(out), C.byref(det))
    return rc, out, fv[:n_vars].copy(), [int(v) for v in un[:out.n_unsatisfied]]


def test_oracle_never_crashes_on_arbitrary_input():
    rng = np.random.default_rng(2026)
    ok = 0
    for k in range(400):
        recs, prios, guesses = arbitrary_setup(rng, wild=k % 2 == 1)
        o = orc.solve(recs, guesses, priorities=prios)
        assert isinstance(o.rc, int)
        ok += o.rc == 0
    assert ok > 20  # some of them are real solves


@pytest.mark.gpu
def test_c_abi_never_crashes_and_agrees_with_the_oracle():
    rng = np.random.default_rng(2026)
    both, compared = 0, 0
    for k in range(400):
        wild = k % 2 == 1
        recs, prios, guesses = arbitrary_setup(rng, wild)
        o = orc.solve(recs, guesses, priorities=prios)
        rc, out, fv, unsat = gpu_solve(recs, prios, guesses)
        assert isinstance(rc, int)
        assert (rc == 0) == (o.rc == 0), (k, rc, o.rc)
        if rc != 0:
            continue
        both += 1
        finite_in = np.isfinite(guesses).all() and np.isfinite(recs["p0"]).all() and np.isfinite(recs["p1"]).all() and np.isfinite(recs["weight"]).all()
        if finite_in and np.isfinite(o.final_values).all() and np.abs(guesses).max(initial=0) < 1e6:
            compared += 1
            assert int(out.iterations) == o.iterations and bool(out.converged) == o.converged, k
            assert int(out.priority_solved) == o.priority_solved and unsat == o.unsatisfied, k
            assert same(fv, o.final_values), (k, fv, o.final_values)
    assert both > 20 and compared > 10
