"""In-kernel phase times of lm_large_kernel (EZPZ_B200_DEBUG=1 prints them per call) and wall time per solve on the
1M-variable chain sketch, a 2D lattice and massive_parallel_system.  usage: EZPZ_B200_DEBUG=1 python tools/time_large.py"""
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ezpz_b200 as ez  # noqa: E402
import workloads as wl  # noqa: E402

ctx = ez.Context(0)
which = sys.argv[1:] or ["chain", "truss", "massive"]
if "chain" in which:
    recs, n, g, _ = wl.chain_sketch(77000)
    t0 = time.perf_counter()
    st = ez.Structure(recs, n)
    t_an = time.perf_counter() - t0
    ts, it, status, path = ctx.time_solve_one(st, g, reps=5)
    print(f"chain sketch n={n}: {statistics.median(ts[1:]) * 1e3:.2f} ms per solve, {it} iterations; host analysis {t_an * 1e3:.0f} ms, "
          f"first solve (device tables uploaded) {ts[0] * 1e3:.0f} ms", flush=True)
    t0 = time.perf_counter()
    st2 = st.extend(recs[-1:])
    print(f"  one constraint added (ezpz_b200_structure_extend, elimination order kept): {(time.perf_counter() - t0) * 1e3:.0f} ms", flush=True)
    del st2
    del st
if "truss" in which:
    for N in (50, 100):
        recs, n, g, _ = wl.grid_truss(N)
        st = ez.Structure(recs, n)
        ts, it, status, path = ctx.time_solve_one(st, g, reps=3)
        print(f"grid_truss({N}) n={n}: {statistics.median(ts[1:]) * 1e3:.2f} ms per solve, {it} iterations", flush=True)
        del st
if "massive" in which:
    recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(500, False))
    st = ez.Structure(recs, n)
    ts, it, status, path = ctx.time_solve_one(st, g, reps=30)
    print(f"massive n={n}: {statistics.median(ts[5:]) * 1e6:.1f} us per solve, {it} iterations", flush=True)
