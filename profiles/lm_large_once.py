"""One solve of a large system through ezpz_b200_solve_one (for ncu captures of lm_large_kernel):
python profiles/lm_large_once.py [cells]   (default 8192 cells = 106,496 variables)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ezpz_b200 as ez  # noqa: E402
import workloads as wl  # noqa: E402

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = ez.Context(0)
recs, n, g, _ = wl.chain_sketch(cells)
st = ez.Structure(recs, n)
out = ctx.solve_one(st, g)
print(n, out.iterations, out.converged, out.path_used)
