"""Device time of the batched kernel on config 2 (65,536 x two_rectangles) or a config-5 fixture; EZPZ_B200_ROLES etc. from the
environment.  usage: python tools/time_small.py [fixture] [batch]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ezpz_b200 as ez  # noqa: E402
import workloads as wl  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "two_rectangles"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
recs, n, g = wl.perturbed_batch(name, B, 0xE2B200D5EED00000)
ctx = ez.Context(0)
st = ez.Structure(recs, n)
dev = torch.device("cuda", 0)
d_g = torch.from_numpy(g).to(dev)
d_f = torch.empty((B, n), dtype=torch.float64, device=dev)
d_it = torch.empty(B, dtype=torch.int32, device=dev)
d_st = torch.empty(B, dtype=torch.uint8, device=dev)
io = {"guesses": d_g.data_ptr(), "final_values": d_f.data_ptr(), "iterations": d_it.data_ptr(), "status": d_st.data_ptr()}
ts = torch.cuda.Stream(device=dev)
for _ in range(5):
    ctx.solve_batch_device(st, io, B, stream=ts.cuda_stream)
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
with torch.cuda.stream(ts):
    for a, b in ev:
        a.record(ts)
        ctx.solve_batch_device(st, io, B, stream=ts.cuda_stream)
        b.record(ts)
torch.cuda.synchronize()
t = sorted(a.elapsed_time(b) for a, b in ev)
it = d_it.cpu().numpy()
print(f"{name} B={B} roles={os.environ.get('EZPZ_B200_ROLES', 'default')}: median {t[len(t) // 2] * 1e3:.1f} us, min {t[0] * 1e3:.1f} us, "
      f"{B / t[len(t) // 2] / 1e3:.1f} M solves/s, iterations mean {it.mean():.2f} max {it.max()}, "
      f"converged {(d_st.cpu().numpy() & 1).mean():.3f}")
