"""ezpz_b200_solve_batch_multi: one call shards a batch over every visible GPU (north_star item 5).  On a
one-GPU box the multi-context has one worker; the same tests run unchanged on 2, 4 or 8 GPUs."""
import numpy as np
import pytest

import ezpz_b200 as ez
import orc
import workloads as wl

pytestmark = pytest.mark.gpu


def _device_count():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def multi():
    return ez.MultiContext()


def test_multi_uses_every_device(multi):
    assert multi.device_count == _device_count() >= 1


@pytest.mark.parametrize("batch", [1, 31, 32, 33, 1000, 20000])
def test_multi_matches_single_device_and_oracle(ctx, multi, batch):
    """Ragged batch sizes (not multiples of 32, fewer problems than devices): every problem bit-identical to the
    single-device call and to the oracle, whatever the number of devices."""
    recs, n, g = wl.two_rectangles_batch(batch)
    st = ez.Structure(recs, n)
    a = multi.solve_batch(st, g, want_unsat=True)
    b = ctx.solve_batch(st, g, want_unsat=True)
    assert np.array_equal(a.final_values.view(np.uint64), b.final_values.view(np.uint64))
    assert np.array_equal(a.iterations, b.iterations) and np.array_equal(a.status, b.status)
    assert np.array_equal(a.unsat_mask, b.unsat_mask)
    k = min(batch, 512)
    fin, it, status = orc.solve_batch(recs, n, g[:k], nthreads=2, hoist=True)
    assert np.array_equal(a.iterations[:k], it) and np.array_equal(a.status[:k] & 3, status & 3)
    assert np.array_equal(a.final_values[:k].view(np.uint64), fin.view(np.uint64))


def test_multi_pinned_buffers_zero_copy_and_pageable_agree(multi):
    """The kernels read and write page-locked caller buffers directly (no staging); pageable buffers take the copy pipeline.
    Same bits either way, verdicts and optional outputs included (mixed structure with unsatisfied problems)."""
    recs, n, g = wl.perturbed_batch("inconsistent", 5000, 0xE2B200D5EED00000 + (5 << 40))
    st = ez.Structure(recs, n)
    pageable = multi.solve_batch(st, g, want_unsat=True, want_degen=True, want_jacobian=True)
    hg, res, owners = ez.pinned_batch_buffers(st, len(g))
    hg[:] = g
    deg = ez.PinnedArray((len(g), st.n_cons), np.uint32)
    jac = ez.PinnedArray((len(g), st.nnz), np.float64)
    res.degen_count, res.jacobian = deg.array, jac.array
    launches = multi.launches
    multi.solve_batch(st, hg, out=res)
    assert multi.launches - launches == multi.device_count  # one kernel per device, no chunking: zero-copy path
    for name in ("final_values", "iterations", "status", "unsat_mask", "degen_count", "jacobian"):
        x, y = getattr(pageable, name), getattr(res, name)
        assert np.array_equal(np.ascontiguousarray(x).view(np.uint8), np.ascontiguousarray(y).view(np.uint8)), name
    assert (res.status & 2).all()  # `inconsistent`: every problem has unsatisfied constraints
    o = orc.solve_inner(recs, g[7])
    bits = np.unpackbits(res.unsat_mask[7].view(np.uint8), bitorder="little")[:st.n_cons]
    assert np.flatnonzero(bits).tolist() == o.unsatisfied == [2, 3, 4, 5]


def test_registered_caller_memory(ctx):
    """ezpz_b200_host_register: what a Rust caller does once with its long-lived Vec<f64> buffers."""
    recs, n, g = wl.two_rectangles_batch(4096)
    st = ez.Structure(recs, n)
    ref = ctx.solve_batch(st, g)
    g2 = g.copy()
    out = ez.BatchResult()
    out.final_values = np.empty_like(g2)
    out.iterations = np.empty(len(g2), np.uint32)
    out.status = np.empty(len(g2), np.uint8)
    out.unsat_mask = np.zeros((len(g2), 1), np.uint32)
    out.degen_count = out.jacobian = None
    arrays = [g2, out.final_values, out.iterations, out.status, out.unsat_mask]
    for a in arrays:
        ez.host_register(a)
    try:
        launches = ctx.launches
        ctx.solve_batch(st, g2, out=out)
        assert ctx.launches - launches == 1
    finally:
        for a in arrays:
            ez.host_unregister(a)
    assert np.array_equal(out.final_values.view(np.uint64), ref.final_values.view(np.uint64))
    assert np.array_equal(out.iterations, ref.iterations) and np.array_equal(out.status, ref.status)


def test_multi_mid_size_systems_and_errors(multi):
    """Structures beyond the thread-per-problem kernel shard the same way (one CTA per problem on each device); an invalid
    call reports instead of hanging the workers."""
    recs, n, g, exact = wl.chain_sketch(16)
    st = ez.Structure(recs, n)
    rng = np.random.default_rng(3)
    G = g[None, :] + rng.uniform(-0.02, 0.02, (9, n))
    out = multi.solve_batch(st, G)
    od = st.ordering()
    for b in (0, 4, 8):
        o = orc.solve_inner_ordered(recs, G[b], od["elim_order"], od["sum_chunk"])
        assert out.iterations[b] == o.iterations
        assert np.array_equal(out.final_values[b].view(np.uint64), o.final_values.view(np.uint64))
    with pytest.raises(ez.EzpzError):
        multi.solve_batch(st, G, params=np.zeros((9, st.n_cons)))  # unsupported on this path: every worker says so
    out2 = multi.solve_batch(st, G)  # the workers are still alive
    assert np.array_equal(out2.final_values, out.final_values)


def test_jobs_of_a_mixed_workload_in_one_call():
    """ezpz_b200_solve_jobs_multi: the eight structure-homogeneous sub-batches of config 5 in ONE call, several workers per
    device (their streams overlap on the GPU); every problem identical to the oracle, verdict masks included."""
    n_dev = _device_count()
    multi = ez.MultiContext(devices=[d for d in range(n_dev) for _ in range(4)])
    assert multi.device_count == 4 * n_dev
    subs = wl.mixed_batches(4096)
    jobs = []
    for name, recs, n, g in subs:
        jobs.append((ez.Structure(recs, n), g))
    results = multi.solve_jobs(jobs, want_unsat=True, want_under=True)
    for (name, recs, n, g), res in zip(subs, results):
        fin, it, status, um, vm = orc.solve_batch(recs, n, g, hoist=True, verdicts=True)
        assert np.array_equal(res.iterations, it), name
        assert np.array_equal(res.final_values.view(np.uint64), fin.view(np.uint64)), name
        assert np.array_equal(res.unsat_mask, um) and np.array_equal(res.under_mask, vm), name
    # a large job is cut over the workers, a failing job reports through its own status without hanging the others
    recs, n, g = wl.two_rectangles_batch(70001)
    st = ez.Structure(recs, n)
    big = multi.solve_jobs([(st, g)])[0]
    ref = multi.solve_batch(st, g)
    assert np.array_equal(big.final_values.view(np.uint64), ref.final_values.view(np.uint64))


def test_two_contexts_on_one_device_share_a_mid_size_structure(ctx):
    """The large path keeps its work buffers per CONTEXT: two workers on the same device solving shards of one mid-size
    structure at the same time (one CTA per problem) must not see each other's state."""
    multi = ez.MultiContext(devices=[0, 0, 0])
    recs, n, g, exact = wl.chain_sketch(16)
    st = ez.Structure(recs, n)
    rng = np.random.default_rng(9)
    G = g[None, :] + rng.uniform(-0.02, 0.02, (600, n))
    ref = ctx.solve_batch(st, G)
    for _ in range(3):
        out = multi.solve_batch(st, G)
        assert np.array_equal(out.final_values.view(np.uint64), ref.final_values.view(np.uint64))
        assert np.array_equal(out.iterations, ref.iterations) and np.array_equal(out.status, ref.status)
