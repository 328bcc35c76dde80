"""Latency of the single-sketch entry points: ezpz_b200_solve_one on an analysed structure, ezpz_b200_solve (priority loop,
lint, topology cache) with and without the cache, next to the CPU port (which re-analyses per solve, as the reference does).
usage: python tools/time_solve_api.py"""
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ezpz_b200 as ez  # noqa: E402
import orc  # noqa: E402
import workloads as wl  # noqa: E402

ctx = ez.Context(0)


def med(fn, reps):
    for _ in range(5):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e6


def raw_solve(cs, analysis):
    """A closure that calls ezpz_b200_solve with prebuilt buffers: the time is the C call's, not the Python wrapper's."""
    import ctypes as C
    native = ez.native
    n_cons, n_vars = len(cs.constraints), len(cs.initial_guesses)
    fv, un, uc = np.zeros(n_vars), np.zeros(n_cons, np.uint64), np.zeros(n_vars, np.uint32)
    warr = (native.WarningRec * (2 * n_cons + 8))()
    out = native.OutcomeRec()
    out.final_values, out.unsatisfied, out.underconstrained = fv.ctypes.data, un.ctypes.data, uc.ctypes.data
    out.warnings, out.warnings_cap = C.addressof(warr), 2 * n_cons + 8
    det = native.ErrorDetail()
    cfg = ez.Config()._native()
    fn = native.lib().ezpz_b200_solve
    args = (ctx.handle, native.ptr(cs.constraints), None, native.ptr(cs.angles_deg), n_cons, None, native.ptr(cs.initial_guesses), n_vars,
            C.byref(cfg), 1 if analysis else 0, C.byref(out), C.byref(det))

    keep = (fv, un, uc, warr, out, det, cfg, cs)  # the C call writes into these: they must outlive raw_solve()

    def go():
        rc = fn(*args)
        assert rc == 0 and keep
    return go


cases = [(name, wl.fixture_text(name)) for name in ("tiny", "two_rectangles", "inconsistent", "nonsquare", "circle_tangent")]
cases += [(f"massive {4 * k}", wl.massive_problem_text(k, False)) for k in (50, 200, 500, 600)]
for name, text in cases:
    recs, n, g, _ = wl.system_from_text(text)
    st = ez.Structure(recs, n)
    cs = ez.textual.Problem(text).to_constraint_system()
    t_one = med(lambda: ctx.solve_one(st, g), 50)
    t_solve = med(raw_solve(cs, False), 50)
    t_ana = med(raw_solve(cs, True), 20 if n < 1000 else 3)
    os.environ["EZPZ_B200_NO_STRUCTURE_CACHE"] = "1"
    ez.native.lib().ezpz_b200_context_clear_cache(ctx.handle)
    t_nocache = med(raw_solve(cs, False), 20)
    os.environ.pop("EZPZ_B200_NO_STRUCTURE_CACHE")
    t_cpu = med(lambda: orc.solve(recs, g), 20)
    t_cpu_ana = med(lambda: orc.solve(recs, g, analysis=True), 20) if n < 1000 else float("nan")
    print(f"{name:16s} n={n:5d}: solve_one {t_one:8.1f} us | ezpz_b200_solve cached {t_solve:8.1f} us, with analysis {t_ana:9.1f} us, "
          f"uncached {t_nocache:8.1f} us | CPU port {t_cpu:8.1f} us, with analysis {t_cpu_ana:10.1f} us")
