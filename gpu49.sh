python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import sys, time
sys.path.insert(0,'tests')
import numpy as np, ezpz_b200 as ez, workloads as wl
ctx = ez.Context(0)
for cells, B in [(16, 4096), (64, 2048), (256, 1024)]:
    recs, n, g, exact = wl.chain_sketch(cells)
    st = ez.Structure(recs, n)
    G = g[None,:] + np.random.default_rng(1).uniform(-0.02,0.02,(B,n))
    out = ctx.solve_batch(st, G)
    t0=time.perf_counter(); out = ctx.solve_batch(st, G); t=time.perf_counter()-t0
    t1=time.perf_counter(); one = ctx.solve_one(st, G[0]); t1=time.perf_counter()-t1
    t1=time.perf_counter(); one = ctx.solve_one(st, G[0]); t1=time.perf_counter()-t1
    print(f"cells {cells} n {n} batch {B}: {t*1e3:.2f} ms  {B/t:.0f} solves/s  (solve_one {t1*1e6:.0f} us -> {1/t1:.0f}/s) iters {out.iterations.min()}..{out.iterations.max()} conv {int((out.status&1).sum())}")
PY
