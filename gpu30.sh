export EZPZ_B200_DEBUG=12
python - <<'PY' 2>&1 | grep -v "^  stage.*tallest"
import sys, time
sys.path.insert(0,'tests')
import ezpz_b200 as ez, workloads as wl
ctx = ez.Context(0)
for build in (lambda: wl.chain_sketch(77000),):
    recs, n, g, _ = build()
    st = ez.Structure(recs, n)
    out = ctx.solve_one(st, g)
PY
