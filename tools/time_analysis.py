"""Host analysis (ezpz_b200_structure_create) of a large sketch, phase by phase: minimum over a few repetitions of the times
EZPZ_B200_DEBUG=1 prints.  usage: python tools/time_analysis.py chain 77000 [reps]   |   truss 100 [reps]
EZPZ_B200_HOST_THREADS=1 gives the single-thread figures."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kind, N, reps = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 5
code = f'''
import sys, time
sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {ROOT!r} + "/tests")
import ezpz_b200 as ez, workloads as wl
recs, n, g, _ = (wl.chain_sketch({N}) if "{kind}" == "chain" else wl.grid_truss({N}))
for i in range({reps}):
    t = time.perf_counter(); st = ez.Structure(recs, n); dt = time.perf_counter() - t
    print(f"[total] analysis {{dt*1e3:.1f}} ms", file=sys.stderr, flush=True)
    del st
'''
env = dict(os.environ, EZPZ_B200_DEBUG="1")
out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stderr
best = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\[(structure|sparse_direct|total)\]\s+(.*?)\s+([0-9.]+) ms$", line)
    if m:
        k = m.group(1) + " " + m.group(2)
        best[k] = min(best.get(k, 1e9), float(m.group(3)))
for k, v in best.items():
    print(f"{k:55s} {v:8.1f} ms")
