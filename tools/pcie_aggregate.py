"""What the HOST side of the box can move: plain cudaMemcpyAsync between page-locked host memory and 1, 2, 4, 8 GPUs at once
(H2D only, D2H only, both directions), per-GPU buffers of the benchmark's size (8.4 MB) and of 64 MB.  No kernels of ours:
this is the ceiling of any host-buffer call, whatever it does on the device.
usage: python tools/pcie_aggregate.py"""
import time

import torch

nd_all = torch.cuda.device_count()
for mb in (8, 64):
    nbytes = mb << 20
    host_in = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(nd_all)]
    host_out = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(nd_all)]
    dev_in = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}") for d in range(nd_all)]
    dev_out = [torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}") for d in range(nd_all)]
    s_in = [torch.cuda.Stream(device=d) for d in range(nd_all)]
    s_out = [torch.cuda.Stream(device=d) for d in range(nd_all)]
    for nd in [d for d in (1, 2, 4, 8) if d <= nd_all]:
        for mode in ("H2D", "D2H", "both"):
            reps = 20

            def go():
                for d in range(nd):
                    if mode in ("H2D", "both"):
                        with torch.cuda.stream(s_in[d]):
                            dev_in[d].copy_(host_in[d], non_blocking=True)
                    if mode in ("D2H", "both"):
                        with torch.cuda.stream(s_out[d]):
                            host_out[d].copy_(dev_out[d], non_blocking=True)
            for _ in range(3):
                go()
            for d in range(nd):
                torch.cuda.synchronize(d)
            t0 = time.perf_counter()
            for _ in range(reps):
                go()
            for d in range(nd):
                torch.cuda.synchronize(d)
            dt = time.perf_counter() - t0
            total = nbytes * nd * reps * (2 if mode == "both" else 1)
            print(f"{mb:3d} MB per GPU and direction, {nd} GPU(s), {mode:4s}: {total / dt / 1e9:7.1f} GB/s aggregate, {total / dt / 1e9 / nd:6.1f} per GPU", flush=True)
