"""Device time of the freedom analysis (freedom_team_kernel) on Jacobians resident in HBM.
usage: python tools/time_freedom.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctypes as C  # noqa: E402

import ezpz_b200 as ez  # noqa: E402
import workloads as wl  # noqa: E402
from ezpz_b200 import native  # noqa: E402

ctx = ez.Context(0)
dev = torch.device("cuda", 0)
ts = torch.cuda.Stream(device=dev)


def time_it(st, jac, reps):
    B = jac.shape[0]
    d_j = torch.from_numpy(np.ascontiguousarray(jac)).to(dev)
    d_m = torch.zeros((B, (st.n_vars + 31) // 32), dtype=torch.int32, device=dev)
    det = native.ErrorDetail()
    fn = native.lib().ezpz_b200_freedom_analysis_device

    def go():
        rc = fn(ctx.handle, st.handle, B, C.c_void_p(d_j.data_ptr()), C.c_void_p(d_m.data_ptr()), C.c_void_p(ts.cuda_stream), C.byref(det))
        assert rc == 0, det.message
    go()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ts):
        a.record(ts)
        for _ in range(reps):
            go()
        b.record(ts)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps, d_m.cpu().numpy()


for name in ("two_rectangles", "underconstrained", "parc_coincident", "square"):
    recs, n, g = wl.perturbed_batch(name, 65536, 0xE2B200D5EED00000)
    st = ez.Structure(recs, n)
    out = ctx.solve_batch(st, g, want_jacobian=True)
    ms, mask = time_it(st, out.jacobian, 5)
    print(f"{name}: 65,536 problems of {st.m} x {n}: {ms * 1e3:.0f} us, {65536 / ms / 1e3:.1f} M analyses/s, {int(mask.any(axis=1).sum())} underconstrained")
for lines in (50, 200, 600, 1100):
    recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(lines, False))
    st = ez.Structure(recs, n)
    one = ctx.solve_one(st, g, want_jacobian=True)
    ms, mask = time_it(st, one.jacobian[None, :], 2)
    print(f"massive {n} x {n}: {ms:.2f} ms")
recs, n, g, exact = wl.chain_sketch(16)
st = ez.Structure(recs, n)
G = g[None, :] + np.random.default_rng(5).uniform(-0.02, 0.02, (256, n))
out = ctx.solve_batch(st, G, want_jacobian=True)
ms, mask = time_it(st, out.jacobian, 2)
print(f"chain sketch 256 problems of {st.m} x {n}: {ms:.2f} ms")
