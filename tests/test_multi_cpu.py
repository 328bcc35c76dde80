"""The N>1 host logic on CPU: two gloo ranks shard a batch with ezpz_b200_shard_range, each 'solves' its
shard with the CPU oracle standing in for the device (this test checks the partition / gather plumbing, not
the kernels), and the gathered result must equal the single-process result problem for problem."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C

    import torch

    import orc
    import workloads as wl
    from ezpz_b200 import native
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B = 1000
    recs, n, g = wl.two_rectangles_batch(B)
    b, e = C.c_uint64(), C.c_uint64()
    native.lib().ezpz_b200_shard_range(B, rank, world, C.byref(b), C.byref(e))
    fin, it, st = orc.solve_batch(recs, n, g[b.value:e.value], nthreads=1, hoist=True)
    # optional final gather of the result SoA: iterations only here
    mine = torch.zeros(B, dtype=torch.int64)
    mine[b.value:e.value] = torch.from_numpy(it.astype(np.int64))
    dist.all_reduce(mine)
    t = torch.tensor([float(e.value - b.value)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        np.save(out_path, mine.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    import orc
    import workloads as wl
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, 29531, out), nprocs=2, join=True)
    gathered = np.load(out)
    recs, n, g = wl.two_rectangles_batch(1000)
    fin, it, st = orc.solve_batch(recs, n, g, nthreads=2, hoist=True)
    assert np.array_equal(gathered, it.astype(np.int64))
