"""e2e time of ezpz_b200_solve_batch (pinned host buffers, 65,536 two_rectangles) against the number of pipeline chunks:
EZPZ_B200_CHUNKS=k python profiles/e2e_chunks.py   (unset = the library's default)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ezpz_b200 as ez  # noqa: E402
from workloads import two_rectangles_batch  # noqa: E402

B = 65536
recs, n, g = two_rectangles_batch(B)
ctx = ez.Context(0)
st = ez.Structure(recs, n)
h_g = torch.from_numpy(g).pin_memory().numpy()
out = ez.BatchResult()
out.final_values = torch.empty((B, n), dtype=torch.float64).pin_memory().numpy()
out.iterations = torch.empty(B, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
out.status = torch.empty(B, dtype=torch.uint8).pin_memory().numpy()
out.unsat_mask = torch.empty((B, (st.n_cons + 31) // 32), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
out.degen_count = None
out.jacobian = None
ts = []
for k in range(25):
    t0 = time.perf_counter()
    ctx.solve_batch(st, h_g, out=out)
    ts.append(time.perf_counter() - t0)
ts = np.array(ts[5:])
assert (out.status & 1).all()
print(f"chunks {os.environ.get('EZPZ_B200_CHUNKS', 'default'):>7s}: median {np.median(ts) * 1e3:.3f} ms  min {ts.min() * 1e3:.3f} ms  "
      f"{B / np.median(ts) / 1e6:.1f} M solves/s")
