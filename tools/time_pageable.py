"""ezpz_b200_solve_batch on ordinary (pageable) numpy arrays — what a Rust Vec<f64> is: staged by the library's host threads
through its page-locked block (default) or left to the driver's staged copies (EZPZ_B200_HOST_MODE=direct), against
page-locked caller buffers.  usage: python tools/time_pageable.py [batch ...]"""
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
for B in [int(a) for a in sys.argv[1:]] or [65536]:
    recs, n, g = wl.perturbed_batch("two_rectangles", B, 0xE2B200D5EED00000)
    st = ez.Structure(recs, n)
    hg, res, owners = ez.pinned_batch_buffers(st, B, want_unsat=True)
    hg[:] = g
    for _ in range(5):
        ctx.solve_batch(st, hg, out=res)
    # caller-owned pageable result buffers, allocated once and reused (a Rust caller's Vecs); `fresh` below lets the
    # Python mirror allocate new result arrays per call instead (first-touch page faults of ~9 MB inside the timed call)
    mine = ez.BatchResult()
    mine.final_values = np.zeros((B, n), np.float64)
    mine.iterations = np.zeros(B, np.uint32)
    mine.status = np.zeros(B, np.uint8)
    mine.unsat_mask = np.zeros((B, (st.n_cons + 31) // 32), np.uint32)
    mine.under_mask = mine.degen_count = mine.jacobian = None
    out = None
    for mode in ("direct", "staged"):
        if mode == "direct":
            os.environ["EZPZ_B200_HOST_MODE"] = "direct"
        else:
            os.environ["EZPZ_B200_HOST_MODE"] = "staged"
        for _ in range(5):
            ctx.solve_batch(st, g, out=mine)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            ctx.solve_batch(st, g, out=mine)
            ts.append(time.perf_counter() - t0)
        ok = np.array_equal(mine.final_values.view(np.uint64), res.final_values.view(np.uint64)) and np.array_equal(mine.iterations, res.iterations)
        print(f"B={B:8d} pageable, buffers reused, {mode:7s}: median {statistics.median(ts) * 1e6:8.1f} us  {B / statistics.median(ts) / 1e6:6.1f} M solves/s  "
              f"results {'identical' if ok else 'DIFFER'}", flush=True)
    for mode in ("direct", "staged"):
        if mode == "direct":
            os.environ["EZPZ_B200_HOST_MODE"] = "direct"
        else:
            os.environ["EZPZ_B200_HOST_MODE"] = "staged"
        for _ in range(5):
            out = ctx.solve_batch(st, g)
        ts = []
        for _ in range(20):
            t0 = time.perf_counter()
            out = ctx.solve_batch(st, g)
            ts.append(time.perf_counter() - t0)
        ok = np.array_equal(out.final_values.view(np.uint64), res.final_values.view(np.uint64)) and np.array_equal(out.iterations, res.iterations)
        print(f"B={B:8d} pageable, fresh results,  {mode:7s}: median {statistics.median(ts) * 1e6:8.1f} us  {B / statistics.median(ts) / 1e6:6.1f} M solves/s  "
              f"results {'identical' if ok else 'DIFFER'}", flush=True)
    os.environ.pop("EZPZ_B200_HOST_MODE", None)
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        ctx.solve_batch(st, hg, out=res)
        ts.append(time.perf_counter() - t0)
    print(f"B={B:8d} page-locked      : median {statistics.median(ts) * 1e6:8.1f} us  {B / statistics.median(ts) / 1e6:6.1f} M solves/s", flush=True)
