python -m pytest tests/test_gpu_large.py -m gpu -x -q 2>&1 | tail -5
export EZPZ_B200_DEBUG=12
python - <<'PY' 2>&1 | grep -v "^  level.*max row"
import sys, time
sys.path.insert(0,'tests')
import ezpz_b200 as ez, workloads as wl
ctx = ez.Context(0)
recs, n, g, _ = wl.chain_sketch(77000)
st = ez.Structure(recs, n)
out = ctx.solve_one(st, g)
PY
