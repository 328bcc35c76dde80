"""Small instances of every kernel path added in round 2, for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_r02.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ezpz_b200 as ez  # noqa: E402
import workloads as wl  # noqa: E402

ctx = ez.Context(0)
# batched kernel with roles + fused freedom analysis (warp-per-problem kernel), zero-copy and staged buffers
for name in ("two_rectangles", "parc_coincident", "underconstrained"):
    recs, n, g = wl.perturbed_batch(name, 300, 0xE2B200D5EED00000)
    st = ez.Structure(recs, n)
    out = ctx.solve_batch(st, g, want_under=True)
    hg, res, owners = ez.pinned_batch_buffers(st, len(g), want_under=True)
    hg[:] = g
    ctx.solve_batch(st, hg, out=res)
    assert np.array_equal(out.under_mask, res.under_mask)
    print(name, "ok", flush=True)
# freedom analysis: CTA per problem on global scratch (208 variables), whole grid on one system (200 variables)
recs, n, g, _ = wl.chain_sketch(16)
st = ez.Structure(recs, n)
G = g[None, :] + np.random.default_rng(1).uniform(-0.02, 0.02, (80, n))
out = ctx.solve_batch(st, G, want_under=True)
print("chain batch ok", int(out.under_mask.any()), flush=True)
recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(50, False))
st = ez.Structure(recs, n)
one = ctx.solve_one(st, g, want_jacobian=True)
print("massive 200 analysis", ctx.freedom_analysis(st, one.jacobian).any(), flush=True)
# large path with row-sliced panels: a 2D lattice on a cluster of 8 CTAs
recs, n, g, _ = wl.grid_truss(24)
st = ez.Structure(recs, n)
o = ctx.solve_one(st, g)
print("grid_truss(24)", o.iterations, o.converged, flush=True)
