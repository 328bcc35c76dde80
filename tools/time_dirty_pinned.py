import os, statistics, sys, time
import numpy as np
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests")
import ezpz_b200 as ez, workloads as wl
ctx = ez.Context(0)
B = 65536
recs, n, g = wl.perturbed_batch("two_rectangles", B, 0xE2B200D5EED00000)
st = ez.Structure(recs, n)
hg, res, owners = ez.pinned_batch_buffers(st, B, want_unsat=True)
hg[:] = g
for _ in range(5): ctx.solve_batch(st, hg, out=res)
for label, pre in (("clean", lambda: None), ("guesses rewritten by the CPU before the call", lambda: hg.__setitem__(slice(None), g)),
                   ("results read by the CPU before the call", lambda: float(res.final_values.sum())),
                   ("both", lambda: (hg.__setitem__(slice(None), g), float(res.final_values.sum())))):
    ts = []
    for _ in range(20):
        pre()
        t0 = time.perf_counter(); ctx.solve_batch(st, hg, out=res); ts.append(time.perf_counter() - t0)
    print(f"page-locked, {label}: median {statistics.median(ts)*1e6:.0f} us", flush=True)
