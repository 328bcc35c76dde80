"""The "underconstrained" verdict (find_dof.rs:15-104) on the device, for any number of variables and in every deployment of
freedom_team_kernel: shared-memory warp/CTA per problem (batches of small sketches), CTA per problem on a global scratch slot
(batches of mid-size sketches), the whole grid on one system (the reference benches its analysis on massive_parallel_system,
solver_bench.rs:146-171, and has no size limit), and fused into the batch solve call (no host round trip in between)."""
import numpy as np
import pytest

import ezpz_b200 as ez
import orc
import workloads as wl

pytestmark = pytest.mark.gpu


def bits(words, count):
    return np.flatnonzero(np.unpackbits(np.ascontiguousarray(words).view(np.uint8), bitorder="little")[:count]).tolist()


def loose_massive(lines, drop_every):
    """massive_parallel_system with the `p_b.y = 4` row of every `drop_every`-th line removed: those lines keep one degree of
    freedom each (the y of their second point), so the system is underconstrained with a nullity of ~lines/drop_every."""
    text = wl.massive_problem_text(lines, False)
    out = []
    for ln in text.splitlines():
        if ln.endswith(".y=4"):
            b = int(ln[1:ln.index(".")])
            if ((b - 1) // 2) % drop_every == 0:
                continue
        out.append(ln)
    return "\n".join(out) + "\n"


@pytest.mark.parametrize("lines", [50, 200])
def test_massive_analysis_matches_oracle(ctx, lines):
    """solver_bench.rs:146-171 runs solve_with_config_analysis on massive_parallel_system (200 variables there); round 1
    refused anything above 256 variables.  Fully constrained: empty list, through the public solve_analysis path."""
    text = wl.massive_problem_text(lines, False)
    out = ez.textual.Problem(text).to_constraint_system().solve_with_config_analysis(ctx=ctx)
    recs, n, g, _ = wl.system_from_text(text)
    o = orc.solve(recs, g, analysis=True)
    assert out.underconstrained == o.underconstrained == []
    assert out.iterations == o.iterations and out.converged


@pytest.mark.parametrize("lines,drop", [(12, 3), (60, 4), (200, 7), (330, 5)])
def test_underconstrained_systems_of_every_size(ctx, lines, drop):
    """48 .. 1,320 variables with a nullity of 4 .. 66: the grid deployment (one system at a time), bit-identical pivoting with
    the oracle, the same underconstrained id list."""
    recs, n, g, _ = wl.system_from_text(loose_massive(lines, drop))
    st = ez.Structure(recs, n)
    one = ctx.solve_one(st, g, want_jacobian=True)
    mask = ctx.freedom_analysis(st, one.jacobian)[0]
    o = orc.solve_inner(recs, g, analysis=True)
    assert one.iterations == o.iterations
    got = bits(mask, n)
    assert got == o.underconstrained
    free = [4 * ln + 3 for ln in range(lines) if ln % drop == 0]  # y of the second point of every loosened line
    assert got == free


def test_massive_2400_variables(ctx):
    """The checked-in massive_parallel_system (600 lines, 2,400 x 2,400) with every 50th line loosened: the analysis the
    reference would run densely.  The expected list is known by construction (the oracle needs ~12 s for this one and is
    run by the test above at 1,320 variables)."""
    lines = 600
    recs, n, g, _ = wl.system_from_text(loose_massive(lines, 50))
    st = ez.Structure(recs, n)
    one = ctx.solve_one(st, g, want_jacobian=True)
    mask = ctx.freedom_analysis(st, one.jacobian)[0]
    assert bits(mask, n) == [4 * ln + 3 for ln in range(lines) if ln % 50 == 0]
    # and through ezpz_b200_solve with analysis (what solve_analysis calls): no EZPZ_ERR_TOO_LARGE any more
    out = ez.textual.Problem(loose_massive(lines, 50)).to_constraint_system().solve_with_config_analysis(ctx=ctx)
    assert out.underconstrained == [4 * ln + 3 for ln in range(lines) if ln % 50 == 0]


def test_more_than_4096_variables(ctx):
    """1,100 lines = 4,400 variables (> 4,096), fully constrained + a few loosened lines."""
    lines = 1100
    recs, n, g, _ = wl.system_from_text(loose_massive(lines, 275))
    st = ez.Structure(recs, n)
    one = ctx.solve_one(st, g, want_jacobian=True)
    mask = ctx.freedom_analysis(st, one.jacobian)[0]
    assert bits(mask, n) == [4 * ln + 3 for ln in range(lines) if ln % 275 == 0]


@pytest.mark.parametrize("name", ["two_rectangles", "underconstrained", "parc_coincident", "perpdist", "parallelogram",
                                  "underdetermined_lines", "arc_radius", "arc_equidistant", "inconsistent"])
def test_fused_analysis_in_the_batch_call(ctx, name):
    """io->under_mask: solve + analysis in one call, Jacobians never leave the device.  Every problem's underconstrained set
    equals the oracle's; the other outputs are bit-identical to a call without the analysis; zero-copy (page-locked) and
    staged (pageable) buffers agree."""
    recs, n, g = wl.perturbed_batch(name, 3000, 0xE2B200D5EED00000 + (11 << 40), half_width=0.05)
    st = ez.Structure(recs, n)
    plain = ctx.solve_batch(st, g, want_unsat=True)
    fused = ctx.solve_batch(st, g, want_unsat=True, want_under=True, want_jacobian=True)
    for k in ("final_values", "iterations", "status", "unsat_mask"):
        assert np.array_equal(np.ascontiguousarray(getattr(plain, k)).view(np.uint8), np.ascontiguousarray(getattr(fused, k)).view(np.uint8)), k
    fin, it, status, um, vm = orc.solve_batch(recs, n, g, hoist=True, verdicts=True)
    assert np.array_equal(fused.iterations, it)
    assert np.array_equal(fused.unsat_mask, um)
    assert np.array_equal(fused.under_mask, vm), (name, np.flatnonzero((fused.under_mask != vm).any(axis=1))[:5])
    # the stand-alone entry on the exported Jacobians gives the same masks
    assert np.array_equal(ctx.freedom_analysis(st, fused.jacobian), vm)
    hg, res, owners = ez.pinned_batch_buffers(st, len(g), want_unsat=True, want_under=True)
    hg[:] = g
    ctx.solve_batch(st, hg, out=res)
    assert np.array_equal(res.under_mask, vm) and np.array_equal(res.final_values.view(np.uint64), fin.view(np.uint64))


def test_fused_analysis_multi_gpu_and_device_pointers(ctx):
    import torch
    multi = ez.MultiContext()
    recs, n, g = wl.perturbed_batch("underconstrained", 10001, 0xE2B200D5EED00000 + (6 << 40))
    st = ez.Structure(recs, n)
    fin, it, status, um, vm = orc.solve_batch(recs, n, g, hoist=True, verdicts=True)
    out = multi.solve_batch(st, g, want_under=True)
    assert np.array_equal(out.under_mask, vm) and vm.any()
    dev = torch.device("cuda", 0)
    B = len(g)
    d_g = torch.from_numpy(g).to(dev)
    d_f = torch.empty((B, n), dtype=torch.float64, device=dev)
    d_it = torch.empty(B, dtype=torch.int32, device=dev)
    d_st = torch.empty(B, dtype=torch.uint8, device=dev)
    d_um = torch.zeros((B, (n + 31) // 32), dtype=torch.int32, device=dev)
    ctx.solve_batch_device(st, {"guesses": d_g.data_ptr(), "final_values": d_f.data_ptr(), "iterations": d_it.data_ptr(),
                                "status": d_st.data_ptr(), "under_mask": d_um.data_ptr()}, B)
    ctx.synchronize()
    assert np.array_equal(d_um.cpu().numpy().view(np.uint32), vm)
    assert np.array_equal(d_f.cpu().numpy().view(np.uint64), fin.view(np.uint64))


def test_batch_of_mid_size_systems(ctx):
    """208-variable sketches, 96 problems: one CTA per problem with the dense matrix in a global scratch slot.  Two structures:
    the full chain (fully constrained) and the chain without two of its CircleRadius rows (those circles keep one degree of
    freedom each: centre y and radius move together along the tangent line)."""
    recs, n, g, exact = wl.chain_sketch(16)
    rng = np.random.default_rng(5)
    G = g[None, :] + rng.uniform(-0.02, 0.02, (96, n))
    for drop_fixed in (False, True):
        r = recs
        if drop_fixed:
            keep = np.ones(len(recs), bool)
            keep[np.flatnonzero(recs["kind"] == 12)[[0, 5]]] = False
            r = np.ascontiguousarray(recs[keep])
        st = ez.Structure(r, n)
        out = ctx.solve_batch(st, G, want_under=True)
        od = st.ordering()
        for b in (0, 17, 95):
            o = orc.solve_inner_ordered(r, G[b], od["elim_order"], od["sum_chunk"])
            assert out.iterations[b] == o.iterations
        # the oracle's analysis at the oracle's own final point (natural order): same verdict
        for b in (0, 95):
            o = orc.solve_inner(r, G[b], analysis=True)
            assert bits(out.under_mask[b], n) == o.underconstrained, drop_fixed
        if drop_fixed:
            assert all(bits(w, n) == [5, 6, 83, 84] for w in out.under_mask)
        else:
            assert not out.under_mask.any()


def test_fused_analysis_through_the_copy_pipeline_under_load():
    """Shards of 32,768 problems and more take the three-stream copy / compute pipeline; the analysis of chunk k reads the
    context's Jacobian buffer while chunk k + 1 is already queued on another stream, and several workers share the device.
    (Regression: the warp-per-problem analysis once skipped the event the next chunk waits for — wrong masks on a loaded
    8-GPU box.)"""
    multi = ez.MultiContext(devices=[0, 0, 0, 0])
    names = ["parc_coincident", "underconstrained", "arc_radius", "perpdist"]
    subs = []
    for k, name in enumerate(names):
        recs, n, g = wl.perturbed_batch(name, 40000, 0xE2B200D5EED00000 + ((20 + k) << 40), half_width=0.05)
        subs.append((name, recs, n, g, ez.Structure(recs, n)))
    want = [orc.solve_batch(recs, n, g, hoist=True, verdicts=True) for name, recs, n, g, st in subs]
    for rep in range(4):
        results = multi.solve_jobs([(st, g) for name, recs, n, g, st in subs], want_unsat=True, want_under=True)
        for (name, recs, n, g, st), res, (fin, it, status, um, vm) in zip(subs, results, want):
            assert np.array_equal(res.under_mask, vm), (name, rep, int((res.under_mask != vm).any(axis=1).sum()))
            assert np.array_equal(res.final_values.view(np.uint64), fin.view(np.uint64)), (name, rep)
