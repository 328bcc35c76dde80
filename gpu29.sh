python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -15
export EZPZ_B200_DEBUG=1
python - <<'PY' 2>&1 | grep -v "^  stage"
import sys, time
sys.path.insert(0,'tests')
import ezpz_b200 as ez, workloads as wl
ctx = ez.Context(0)
for build in (lambda: wl.system_from_text(wl.massive_problem_text(500, False)), lambda: wl.chain_sketch(1024), lambda: wl.chain_sketch(8192), lambda: wl.chain_sketch(77000)):
    recs, n, g, _ = build()
    st = ez.Structure(recs, n)
    for k in range(2):
        t=time.perf_counter(); out = ctx.solve_one(st, g); dt=time.perf_counter()-t
        print("n", n, "solve_one %.1f us"%(dt*1e6), out.iterations, out.path_used)
PY
