"""Small instances of the kernel paths added late in round 2, for compute-sanitizer (memcheck / racecheck):
   EZPZ_B200_PIPE_ASM=2 compute-sanitizer --tool memcheck python tools/sanitize_r02b.py
 * the assembly phases inside lm_large_kernel through the per-warp cp.async.bulk pipeline (assemble_phase_pipe), forced on
   a single-CTA system, a cluster of eight CTAs and a whole-grid launch;
 * the lane form of the host-buffer batch pipeline (one copy-in stream, two kernel streams, one copy-out stream);
 * a solve through an extended structure (ezpz_b200_structure_extend) whose device tables are copied table by table."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ezpz_b200 as ez  # noqa: E402
import workloads as wl  # noqa: E402

ctx = ez.Context(0)
for cells in (8, 64, 1024):
    recs, n, g, _ = wl.chain_sketch(cells)
    st = ez.Structure(recs, n)
    o = ctx.solve_one(st, g)
    print(f"chain_sketch({cells}) n={n}", o.iterations, o.converged, flush=True)
recs, n, g, _ = wl.grid_truss(24)
base = ez.Structure(recs[:-5], n)
st = base.extend(recs[-5:])
o = ctx.solve_one(st, g)
print("grid_truss(24) through an extended structure", o.iterations, o.converged, flush=True)
os.environ["EZPZ_B200_HOST_MODE"] = "pipeline"
os.environ["EZPZ_B200_PIPE_LANES"] = "6"
recs, n, g = wl.perturbed_batch("two_rectangles", 16384 + 37, 0xE2B200D5EED00000)
st = ez.Structure(recs, n)
ref = ctx.solve_batch(st, g)
hg, res, owners = ez.pinned_batch_buffers(st, len(g), want_unsat=True)
hg[:] = g
ctx.solve_batch(st, hg, out=res)
assert np.array_equal(ref.final_values.view(np.uint64), res.final_values.view(np.uint64)) and np.array_equal(ref.iterations, res.iterations)
print("lane pipeline ok", flush=True)
