python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY' 2>&1 | tail -12
import sys, os, time
sys.path.insert(0,'tests')
import numpy as np, ezpz_b200 as ez, workloads as wl
ctx = ez.Context(0)
for lines in (500, 600):
    recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(lines, False))
    st = ez.Structure(recs, n)
    for k in range(5): out = ctx.solve_one(st, g)
    ts=[]
    for k in range(200):
        t=time.perf_counter(); out = ctx.solve_one(st, g); ts.append(time.perf_counter()-t)
    print(f"massive {n}x{n} solve_one us: median {np.median(ts)*1e6:.1f} mean {np.mean(ts)*1e6:.1f} iters {out.iterations}")
os.environ["EZPZ_B200_DEBUG"]="1"
recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(500, False))
st = ez.Structure(recs, n)
for k in range(3): out = ctx.solve_one(st, g)
PY
