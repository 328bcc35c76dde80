EZPZ_B200_DEBUG=12 python profiles/lm_large_once.py 77000 2>&1 | grep -v "^\[sparse_direct\] [a-z]" | tail -24
python - <<'PY' 2>&1 | tail -8
import sys, os
sys.path.insert(0,'tests')
os.environ["EZPZ_B200_DEBUG"]="1"
import numpy as np, ezpz_b200 as ez, workloads as wl
ctx = ez.Context(0)
recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(500, False))
st = ez.Structure(recs, n)
for k in range(3): out = ctx.solve_one(st, g)
PY
