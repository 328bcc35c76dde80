"""Where the host-buffer batch call spends its time: lm_small_kernel timed with CUDA events for every combination of guesses /
results living in HBM or in page-locked host memory (read / written by the kernel across PCIe).
usage: python tools/time_e2e_split.py [batch]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ezpz_b200 as ez  # noqa: E402
import workloads as wl  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
recs, n, g = wl.perturbed_batch("two_rectangles", B, 0xE2B200D5EED00000)
ctx = ez.Context(0)
st = ez.Structure(recs, n)
dev = torch.device("cuda", 0)
d_g = torch.from_numpy(g).to(dev)
d_f = torch.empty((B, n), dtype=torch.float64, device=dev)
d_it = torch.empty(B, dtype=torch.int32, device=dev)
d_st = torch.empty(B, dtype=torch.uint8, device=dev)
d_un = torch.empty(B, dtype=torch.int32, device=dev)
hg, res, owners = ez.pinned_batch_buffers(st, B, want_unsat=True)
hg[:] = g
ts = torch.cuda.Stream(device=dev)


def run(name, io):
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
    print(f"{name:46s} median {t[len(t) // 2] * 1e3:7.1f} us  min {t[0] * 1e3:7.1f} us")


dev_in = {"guesses": d_g.data_ptr()}
host_in = {"guesses": hg.ctypes.data}
dev_out = {"final_values": d_f.data_ptr(), "iterations": d_it.data_ptr(), "status": d_st.data_ptr(), "unsat_mask": d_un.data_ptr()}
host_out = {"final_values": res.final_values.ctypes.data, "iterations": res.iterations.ctypes.data, "status": res.status.ctypes.data,
            "unsat_mask": res.unsat_mask.ctypes.data}
run("guesses HBM,  results HBM", {**dev_in, **dev_out})
run("guesses host, results HBM", {**host_in, **dev_out})
run("guesses HBM,  results host", {**dev_in, **host_out})
run("guesses host, results host (the zero-copy call)", {**host_in, **host_out})
# plain DMA of the same bytes for reference
hgt = torch.from_numpy(g).pin_memory()
hft = torch.empty((B, n), dtype=torch.float64).pin_memory()
for name, fn in (("cudaMemcpyAsync H2D of the guesses", lambda: d_g.copy_(hgt, non_blocking=True)),
                 ("cudaMemcpyAsync D2H of the finals", lambda: hft.copy_(d_f, non_blocking=True))):
    with torch.cuda.stream(ts):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(ts)
        for _ in range(10):
            fn()
        b.record(ts)
    torch.cuda.synchronize()
    print(f"{name:46s} {a.elapsed_time(b) / 10 * 1e3:7.1f} us ({B * n * 8 / (a.elapsed_time(b) / 10 * 1e-3) / 1e9:.1f} GB/s)")
