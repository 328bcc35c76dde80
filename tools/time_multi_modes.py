"""ONE ezpz_b200_solve_batch_multi call per step over 1, 2, 4, 8 GPUs on page-locked caller buffers, in the two host-buffer
forms (EZPZ_B200_HOST_MODE).  Weak (65,536 problems per GPU) and strong (65,536 in total) batches.  Wall clock per call.
usage: python tools/time_multi_modes.py"""
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ezpz_b200 as ez  # noqa: E402
import workloads as wl  # noqa: E402

ndev_all = torch.cuda.device_count()
PER = 65536
recs, n, g_all = wl.perturbed_batch("two_rectangles", PER * ndev_all, 0xE2B200D5EED00000)
st = ez.Structure(recs, n)
hg, res, owners = ez.pinned_batch_buffers(st, PER * ndev_all, want_unsat=True)
hg[:] = g_all
base = {}
for nd in [d for d in (1, 2, 4, 8) if d <= ndev_all]:
    multi = ez.MultiContext(devices=list(range(nd)))
    for label, B in (("weak", PER * nd), ("strong", PER)):
        for mode in ("zerocopy", "pipeline"):
            os.environ["EZPZ_B200_HOST_MODE"] = mode
            sub = ez.BatchResult()
            sub.final_values, sub.iterations, sub.status, sub.unsat_mask = res.final_values[:B], res.iterations[:B], res.status[:B], res.unsat_mask[:B]
            sub.degen_count = sub.jacobian = sub.under_mask = None
            for _ in range(10):
                multi.solve_batch(st, hg[:B], out=sub)
            ts = []
            for _ in range(30):
                t0 = time.perf_counter()
                multi.solve_batch(st, hg[:B], out=sub)
                ts.append(time.perf_counter() - t0)
            med = statistics.median(ts)
            key = (label, mode)
            if nd == 1:
                base[key] = B / med
            eff = (B / med) / (base[key] * (nd if label == "weak" else 1))
            print(f"{nd} GPU {label:6s} B={B:7d} {mode:9s} median {med * 1e6:7.1f} us  min {min(ts) * 1e6:7.1f} us  {B / med / 1e6:7.1f} M solves/s"
                  + (f"  efficiency {eff:.2f}" if label == "weak" else f"  speed-up {eff:.2f}"), flush=True)
    del multi
