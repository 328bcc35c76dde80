"""BASELINE.json config 5: batch-size sweep 1K-1M of MIXED sketches (50 % two_rectangles, 15 % square, 10 % circle_tangent,
5 % each arc_length, parc_coincident (underconstrained), inconsistent, underconstrained, perpendicular), grouped into
structure-homogeneous sub-batches, sharded over the ranks (one process per GPU, no data-path collective).

    python profiles/sweep_mixed.py                       # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/sweep_mixed.py

Per total batch size: device time (CUDA events around all sub-batch launches of the rank's shard, guesses resident in HBM,
max over ranks) -> solves/s; verdict parity of a sample of every sub-batch against the CPU oracle (iterations, converged /
unsatisfied status, coordinates bit for bit).  One JSON line per batch size on rank 0."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ezpz_b200 as ez  # noqa: E402
import orc  # noqa: E402
import workloads as wl  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = ez.Context(local)
    sizes = [1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20]
    if len(sys.argv) > 1:
        sizes = [int(a) for a in sys.argv[1:]]
    tstream = torch.cuda.Stream(device=dev)
    structures = {}
    sub_streams = []
    for total in sizes:
        subs = []
        for name, recs, n, g in wl.mixed_batches(total):
            b, e = ez.shard_range(len(g), rank, world)
            if name not in structures:
                structures[name] = ez.Structure(recs, n)
            st = structures[name]
            cnt = e - b
            d_g = torch.from_numpy(np.ascontiguousarray(g[b:e])).to(dev)
            d_f = torch.empty((max(cnt, 1), n), dtype=torch.float64, device=dev)
            d_it = torch.empty(max(cnt, 1), dtype=torch.int32, device=dev)
            d_st = torch.empty(max(cnt, 1), dtype=torch.uint8, device=dev)
            io = {"guesses": d_g.data_ptr(), "final_values": d_f.data_ptr(), "iterations": d_it.data_ptr(),
                  "status": d_st.data_ptr()}
            subs.append((name, recs, n, g[b:e], st, cnt, io, d_g, d_f, d_it, d_st))

        # one stream per structure: the eight sub-batch kernels of a pass are independent and overlap when a
        # sub-batch is too small to fill the GPU (the long-running ones are the inconsistent systems: 35 iterations)
        while len(sub_streams) < len(subs):
            sub_streams.append(torch.cuda.Stream(device=dev))

        def step():
            fork = torch.cuda.Event()
            fork.record(tstream)
            for k, (name, recs, n, g, st, cnt, io, *_) in enumerate(subs):
                if cnt:
                    sub_streams[k].wait_event(fork)
                    ctx.solve_batch_device(st, io, cnt, stream=sub_streams[k].cuda_stream)
                    join = torch.cuda.Event()
                    join.record(sub_streams[k])
                    tstream.wait_event(join)

        with torch.cuda.stream(tstream):
            for _ in range(3):
                step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(tstream):
            e0.record(tstream)
            for _ in range(reps):
                step()
            e1.record(tstream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # verdict parity on a sample of every sub-batch of this rank's shard
        bad = 0
        verdicts = {}
        for name, recs, n, g, st, cnt, io, d_g, d_f, d_it, d_st in subs:
            if not cnt:
                continue
            k = min(cnt, 256)
            fin, it, status = orc.solve_batch(recs, n, g[:k], hoist=True)
            gf, gi, gs = d_f[:k].cpu().numpy(), d_it[:k].cpu().numpy().view(np.uint32), d_st[:k].cpu().numpy()
            bad += int((gi != it).sum()) + int(((gs & 3) != (status & 3)).sum()) + int((gf.view(np.uint64) != fin.view(np.uint64)).any(axis=1).sum())
            s_all = d_st[:cnt].cpu().numpy()
            verdicts[name] = {"converged": int((s_all & 1).sum()), "unsatisfied": int(((s_all >> 1) & 1).sum()), "of": int(cnt)}
        t = torch.tensor([bad], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        if rank == 0:
            n_total = sum(len(x[3]) for x in subs) if world == 1 else sum(len(m[3]) for m in wl.mixed_batches(total))
            print(json.dumps({"config": "mixed sketches (BASELINE.json configs[4])", "n_gpus": world, "batch_total": n_total,
                              "ms_per_pass": float(ms.item()), "solves_per_s": n_total / (float(ms.item()) * 1e-3),
                              "launches_per_pass": sum(1 for x in subs if x[5]),
                              "parity_mismatches_in_sample": int(t.item()), "rank0_verdicts": verdicts}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
