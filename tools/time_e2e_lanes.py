"""Host-buffer batch call, pipeline form: the three-stream form (EZPZ_B200_PIPE_LANES=0) against the lane form with 2..16 chunks.
Wall clock per call on page-locked buffers.  usage: python tools/time_e2e_lanes.py [batch ...]"""
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ezpz_b200 as ez  # noqa: E402
import workloads as wl  # noqa: E402

ctx = ez.Context(0)
os.environ["EZPZ_B200_HOST_MODE"] = "pipeline"
for B in [int(a) for a in sys.argv[1:]] or [65536]:
    recs, n, g = wl.perturbed_batch("two_rectangles", B, 0xE2B200D5EED00000)
    st = ez.Structure(recs, n)
    hg, res, owners = ez.pinned_batch_buffers(st, B, want_unsat=True)
    hg[:] = g
    ref = ctx.solve_batch(st, g)
    for lanes in (0, 2, 3, 4, 5, 6, 8, 10, 12, 16):
        os.environ["EZPZ_B200_PIPE_LANES"] = str(lanes)
        res.final_values[:] = 0
        for _ in range(10):
            ctx.solve_batch(st, hg, out=res)
        ok = np.array_equal(res.final_values.view(np.uint64), ref.final_values.view(np.uint64)) and np.array_equal(res.iterations, ref.iterations) \
            and np.array_equal(res.status, ref.status) and np.array_equal(res.unsat_mask, ref.unsat_mask)
        ts = []
        for _ in range(40):
            t0 = time.perf_counter()
            ctx.solve_batch(st, hg, out=res)
            ts.append(time.perf_counter() - t0)
        print(f"B={B:8d} lanes {lanes:2d}  median {statistics.median(ts) * 1e6:7.1f} us  min {min(ts) * 1e6:7.1f} us  "
              f"{B / statistics.median(ts) / 1e6:7.1f} M solves/s  results {'identical' if ok else 'DIFFER'}", flush=True)
