import os, sys, time
ROOT = "/root/repo" if os.path.exists("/root/repo/bench.py") else os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import ezpz_b200 as ez
from workloads import two_rectangles_batch
B = 65536
recs, n, g = two_rectangles_batch(B)
ctx = ez.Context(0); st = ez.Structure(recs, n)
h_g = torch.from_numpy(g).pin_memory().numpy()
out = ez.BatchResult()
out.final_values = torch.empty((B, n), dtype=torch.float64).pin_memory().numpy()
out.iterations = torch.empty(B, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
out.status = torch.empty(B, dtype=torch.uint8).pin_memory().numpy()
out.unsat_mask = torch.empty((B, (st.n_cons + 31) // 32), dtype=torch.int32).pin_memory().numpy().view(np.uint32)
out.degen_count = None; out.jacobian = None
ts = []
for k in range(40):
    t0 = time.perf_counter(); ctx.solve_batch(st, h_g, out=out); ts.append((time.perf_counter() - t0) * 1e3)
print("e2e ms per call:", " ".join(f"{t:.2f}" for t in ts))
