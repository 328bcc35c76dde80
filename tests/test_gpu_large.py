"""GPU parity for single large systems (configs 3 and 4 of BASELINE.json) through ezpz_b200_solve_one."""
import os

import numpy as np
import pytest

import ezpz_b200 as ez
import orc
import workloads as wl
from test_gpu_parity import assert_bitwise

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lines,over", [(500, False), (600, False), (500, True)])
def test_massive_parallel_system_direct_path(ctx, lines, over):
    """Config 3: 2,000 x 2,000 (README size), the checked-in 2,400 x 2,400 and the overconstrained variant.
    Sparse direct solve in natural order: bit-exact against the oracle (same sum-of-squares chunking)."""
    recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(lines, over))
    st = ez.Structure(recs, n)
    out = ctx.solve_one(st, g)
    assert out.path_used == 1
    od = st.ordering()
    assert not od["nested"] and od["sum_chunk"] == 64  # natural order; one cluster of 8 CTAs, chunked sum of squares
    o = orc.solve_inner_ordered(recs, g, None, od["sum_chunk"])
    ref = orc.solve_inner(recs, g)  # the reference-faithful sequential sum: same trajectory, same solution to 1e-9
    assert ref.iterations == o.iterations and np.abs(ref.final_values - o.final_values).max() <= 1e-9
    assert out.iterations == o.iterations and out.converged and out.unsatisfied == []
    if not over:
        assert out.iterations == 2  # README.md:38 "Iterations needed: 2"
    assert_bitwise(out.final_values, o.final_values, "final values")
    for k in range(lines):  # every line ends up vertical at x = k, from y = 0 to y = 4
        assert np.abs(out.final_values[4 * k:4 * k + 4] - [k, 0, k, 4]).max() < 1e-6


def _check_direct(ctx, cells, system=None, exact_tol=1e-6):
    """Sparse direct path: bit-exact against the oracle run with the same elimination order and sum-of-squares
    chunking (Structure.ordering()), and within the north-star tolerance (identical iteration count and verdict,
    1e-9 on coordinates) of the reference-faithful oracle (natural order, sequential sum)."""
    recs, n, g, exact = system if system is not None else wl.chain_sketch(cells)
    st = ez.Structure(recs, n)
    od = st.ordering()
    assert od["path"] == 1 and sorted(od["elim_order"].tolist()) == list(range(n))
    out = ctx.solve_one(st, g)
    assert out.path_used == 1
    o = orc.solve_inner_ordered(recs, g, od["elim_order"], od["sum_chunk"])
    assert out.iterations == o.iterations and out.converged == o.converged and out.unsatisfied == o.unsatisfied
    assert_bitwise(out.final_values, o.final_values, "final values")
    ref = orc.solve_inner(recs, g)
    assert out.iterations == ref.iterations and out.converged == ref.converged and out.unsatisfied == ref.unsatisfied
    scale = np.maximum(1.0, np.abs(ref.final_values))
    assert (np.abs(out.final_values - ref.final_values) <= 1e-9 * scale).all()
    assert np.abs(out.final_values - exact).max() < exact_tol  # the constructed solution (LM stops at max|r| <= 1e-8)
    return od


@pytest.mark.parametrize("cells", [8, 16])
def test_chain_sketch_direct_single_cta(ctx, cells):
    od = _check_direct(ctx, cells)  # 104 / 208 variables: one CTA, sequential sum of squares
    assert od["path"] == 1 and od["sum_chunk"] == 0


def test_chain_sketch_direct_cluster(ctx):
    od = _check_direct(ctx, 64)  # 832 variables: one thread-block cluster of 8 CTAs, nested-dissection order
    assert od["nested"] and od["sum_chunk"] == 64


@pytest.mark.parametrize("cells", [1024, 8192])
def test_chain_sketch_direct_grid(ctx, cells):
    od = _check_direct(ctx, cells)  # cooperative grid, levels run grid-wide then by one CTA
    assert od["nested"] and od["sum_chunk"] == (128 if cells == 1024 else 512)  # the power of two next to sqrt(m)


def test_chain_sketch_1m_variables(ctx):
    """BASELINE.json config 4: the synthetic 1,001,000-variable sketch (arcs, circle tangents, distances, angles)."""
    od = _check_direct(ctx, 77000)
    assert od["n_levels"] < 1024


@pytest.mark.parametrize("N,weights", [(12, False), (40, False), (40, True), (100, False)])
def test_grid_truss_direct(ctx, N, weights):
    """A 2D lattice (separators of ~N points, panels hundreds of rows tall, dozens of updates per panel): the regime
    opposite to the chain sketch — panels that do not fit a warp's stage, CTA teams, column-sliced update blocks."""
    od = _check_direct(ctx, 0, wl.grid_truss(N, weights=weights), exact_tol=1e-4)
    assert od["nested"]


@pytest.mark.parametrize("system", ["chain", "truss"])
def test_extended_structure_solves_like_the_oracle_in_the_kept_order(ctx, system):
    """ezpz_b200_structure_extend on the sparse direct path: constraints added to an analysed sketch, the base's elimination
    order kept; the solve through the extended structure is bit-exact against the oracle run in that order and agrees with a
    freshly analysed structure's solve (identical iterations and verdict, 1e-9)."""
    recs, n, g, exact = wl.chain_sketch(1024) if system == "chain" else wl.grid_truss(30)
    k = len(recs) - 11
    base = ez.Structure(recs[:k], n)
    ext = base.extend(recs[k:])
    od = ext.ordering()
    assert od["path"] == 1 and np.array_equal(od["elim_order"], base.ordering()["elim_order"])
    out = ctx.solve_one(ext, g)
    o = orc.solve_inner_ordered(recs, g, od["elim_order"], od["sum_chunk"])
    assert out.iterations == o.iterations and out.converged == o.converged and out.unsatisfied == o.unsatisfied
    assert_bitwise(out.final_values, o.final_values, "final values")
    fresh = ctx.solve_one(ez.Structure(recs, n), g)
    assert out.iterations == fresh.iterations and out.converged == fresh.converged and out.unsatisfied == fresh.unsatisfied
    scale = np.maximum(1.0, np.abs(fresh.final_values))
    assert (np.abs(out.final_values - fresh.final_values) <= 1e-9 * scale).all()
    # repeats of existing constraints: no new coupling, the base's whole schedule is taken over (only the products of A = JtJ
    # are rebuilt); still the oracle's arithmetic on the longer list in the kept order
    again = np.concatenate([recs[3:5], recs[k:k + 2]])
    ext2 = ext.extend(again)
    recs2 = np.concatenate([recs, again])
    od2 = ext2.ordering()
    assert np.array_equal(od2["elim_order"], od["elim_order"])
    out2 = ctx.solve_one(ext2, g)
    o2 = orc.solve_inner_ordered(recs2, g, od2["elim_order"], od2["sum_chunk"])
    assert out2.iterations == o2.iterations and out2.converged == o2.converged and out2.unsatisfied == o2.unsatisfied
    assert_bitwise(out2.final_values, o2.final_values, "final values")


@pytest.mark.parametrize("system", ["chain104", "chain13312", "truss40", "massive"])
def test_pipelined_assembly_inside_the_lm_kernel(ctx, system, monkeypatch):
    """The assembly phases inside lm_large_kernel through the per-warp cp.async.bulk tile pipeline (assemble_phase_pipe; by
    default only when every warp has at least four record tiles, i.e. beyond ~300,000 constraints) forced on systems of every
    launch shape — one CTA, a cluster of eight, the whole grid: the same bits as the oracle."""
    monkeypatch.setenv("EZPZ_B200_PIPE_ASM", "2")
    if system == "massive":
        recs, n, g, _ = wl.system_from_text(wl.massive_problem_text(500, False))
        st = ez.Structure(recs, n)
        od = st.ordering()
        out = ctx.solve_one(st, g)
        o = orc.solve_inner_ordered(recs, g, None, od["sum_chunk"])
        assert out.iterations == o.iterations == 2 and out.converged
        assert_bitwise(out.final_values, o.final_values, "final values")
    elif system == "truss40":
        _check_direct(ctx, 0, wl.grid_truss(40), exact_tol=1e-4)
    else:
        _check_direct(ctx, int(system[5:]) // 13)


def test_chain_sketch_pcg_path(ctx):
    """PCG path forced on a 13,312-variable sketch: the step is solved iteratively to 1e-13 relative residual,
    so results agree with the oracle's direct solve to 1e-9 and the LM trajectory has the same length."""
    os.environ["EZPZ_B200_FORCE_PCG"] = "1"
    try:
        recs, n, g, exact = wl.chain_sketch(1024)
        st = ez.Structure(recs, n)
    finally:
        del os.environ["EZPZ_B200_FORCE_PCG"]
    out = ctx.solve_one(st, g)
    assert out.path_used == 2 and out.lin_iters > 0
    o = orc.solve_inner(recs, g)
    assert out.converged and out.unsatisfied == []
    assert out.iterations == o.iterations
    scale = np.maximum(1.0, np.abs(o.final_values))
    assert (np.abs(out.final_values - o.final_values) <= 1e-9 * scale).all()


def test_large_eval_matches_oracle_bitwise(ctx):
    """Assembly kernel on a 13k-variable system: residuals and Jacobian values bit for bit."""
    from test_gpu_parity import resolved
    recs, n, g, exact = wl.chain_sketch(1024)
    st = ez.Structure(recs, n)
    r, jc, jr, dg = ctx.evaluate(st, g)
    ro, jo, dgo, _ = orc.evaluate(resolved(recs, g), n, g)
    assert_bitwise(r, ro, "residual")
    assert_bitwise(jc, jo, "jacobian")


def test_large_eval_all_kinds_bitwise(ctx):
    """All 25 kinds (with weights, undefined and given tangent sides, degenerate geometry) through the large
    path's record tiles and bulk-copy-staged assembly kernel, in both record layouts (direct: CSC only; PCG: CSC +
    CSR-ordered copy), bit for bit against the oracle."""
    from test_gpu_parity import random_constraints, resolved
    rng = np.random.default_rng(20261018)
    for trial in range(6):
        n_vars = 400
        cons = random_constraints(rng, 1500, n_vars)
        weights = np.ones(len(cons)) if trial % 2 == 0 else rng.choice([1.0, 0.5, 3.0], len(cons))
        recs = ez.records(cons, weights)
        x = rng.uniform(-8.0, 8.0, n_vars)
        if trial >= 4:  # collapse some points onto each other
            x[2:4] = x[0:2]
            x[8:10] = x[6:8]
        if trial % 3 == 2:
            os.environ["EZPZ_B200_FORCE_PCG"] = "1"
        try:
            st = ez.Structure(recs, n_vars)
        finally:
            os.environ.pop("EZPZ_B200_FORCE_PCG", None)
        assert st.ordering()["path"] == (2 if trial % 3 == 2 else 1)
        r, jc, jr, dg = ctx.evaluate(st, x)
        ro, jo, dgo, _ = orc.evaluate(resolved(recs, x), n_vars, x)
        assert_bitwise(r, ro, f"trial {trial} residual")
        assert_bitwise(jc, jo, f"trial {trial} jacobian (CSC)")
        pat = st.pattern()
        order = np.lexsort((np.repeat(np.arange(n_vars), np.diff(pat["csc_col_ptr"])), pat["csc_row_idx"]))
        assert_bitwise(jr, jo[order], f"trial {trial} jacobian (CSR order)")
        assert np.array_equal(dg, dgo)


@pytest.mark.parametrize("cells", [64, 1024])
def test_large_path_failed_factorisations(ctx, cells):
    """Non-finite pivots (a guess at 1e308 overflows the distances): every factorisation fails, lambda is bumped and the
    iteration is burnt (newton.rs:96-99) until max_iterations — same trajectory length and verdict as the oracle, on the
    single-CTA and on the cooperative-grid variant."""
    recs, n, g, exact = wl.chain_sketch(cells)
    g = g.copy()
    g[2] = 1e308
    st = ez.Structure(recs, n)
    od = st.ordering()
    out = ctx.solve_one(st, g)
    o = orc.solve_inner_ordered(recs, g, od["elim_order"], od["sum_chunk"])
    assert out.iterations == o.iterations == 35 and not out.converged and not o.converged
    assert out.unsatisfied == o.unsatisfied
    finite = np.isfinite(o.final_values)
    assert np.array_equal(np.isfinite(out.final_values), finite)
    assert_bitwise(out.final_values[finite], o.final_values[finite], "finite final values")


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_systems_all_kinds_large_path(ctx, seed):
    """Random constraint soups (all 25 kinds, weights, given and undefined tangent sides; inconsistent, rank deficient,
    sometimes degenerate) on 400 variables through the large path: whatever the LM trajectory does — failed
    factorisations, rejected steps, no convergence — it is the oracle's trajectory (same order and sum chunking):
    iteration count, verdict bits and every finite coordinate bit for bit."""
    from test_gpu_parity import random_constraints
    rng = np.random.default_rng(1000 + seed)
    n_vars = 400
    cons = random_constraints(rng, 420 + 40 * seed, n_vars)
    weights = rng.choice([1.0, 0.5, 3.0], len(cons)) if seed % 2 else np.ones(len(cons))
    recs = ez.records(cons, weights)
    g = rng.uniform(-8.0, 8.0, n_vars)
    st = ez.Structure(recs, n_vars)
    od = st.ordering()
    assert od["path"] == 1
    out = ctx.solve_one(st, g)
    o = orc.solve_inner_ordered(recs, g, od["elim_order"], od["sum_chunk"])
    assert out.iterations == o.iterations and out.converged == o.converged
    assert out.unsatisfied == o.unsatisfied
    finite = np.isfinite(o.final_values)
    assert np.array_equal(np.isfinite(out.final_values), finite)
    assert_bitwise(out.final_values[finite], o.final_values[finite], "finite final values")


@pytest.mark.parametrize("cells,batch", [(16, 40), (64, 24)])
def test_batch_of_mid_size_systems_one_cta_per_problem(ctx, cells, batch):
    """ezpz_b200_solve_batch on a structure beyond the thread-per-problem kernel (208 / 832 variables): the persistent LM
    kernel with one CTA per problem.  Every problem is bit-identical to solving it alone (ezpz_b200_solve_one: one CTA
    resp. one cluster of 8 CTAs) and to the oracle with the structure's order and sum chunking; degenerate counts, the
    unsatisfied mask and the exported Jacobian included.  One problem carries a 1e308 guess (failed factorisations)."""
    recs, n, g, exact = wl.chain_sketch(cells)
    st = ez.Structure(recs, n)
    od = st.ordering()
    rng = np.random.default_rng(77 + cells)
    G = g[None, :] + rng.uniform(-0.02, 0.02, (batch, n))
    G[0] = g
    G[3, 2] = 1e308
    os.environ["EZPZ_B200_LARGE_BATCH_CAP"] = "16"  # several launches per call
    try:
        out = ctx.solve_batch(st, G, want_unsat=True, want_degen=True, want_jacobian=True)
    finally:
        del os.environ["EZPZ_B200_LARGE_BATCH_CAP"]
    for b in range(batch):
        one = ctx.solve_one(st, G[b], want_jacobian=True)
        o = orc.solve_inner_ordered(recs, G[b], od["elim_order"], od["sum_chunk"])
        assert out.iterations[b] == one.iterations == o.iterations, b
        assert (out.status[b] & 3) == (one.status & 3) and bool(out.status[b] & 1) == o.converged, b
        finite = np.isfinite(o.final_values)
        assert_bitwise(out.final_values[b][finite], o.final_values[finite], f"problem {b} final values vs oracle")
        assert np.array_equal(out.final_values[b], one.final_values, equal_nan=True), b
        assert np.array_equal(out.unsat_mask[b], one.unsat_mask), b
        assert np.array_equal(out.degen_count[b], one.degen_count), b
        assert np.array_equal(out.jacobian[b], one.jacobian, equal_nan=True), b
    assert out.iterations[3] == 35 and not (out.status[3] & 1)


def test_batch_of_mid_size_systems_pcg_path(ctx):
    """The one-CTA-per-problem batch on the PCG path (forced): per-problem CSR copies and CG vectors; every problem equals the
    same problem solved alone, bit for bit (same kernel code, same reduction shapes inside one CTA)."""
    os.environ["EZPZ_B200_FORCE_PCG"] = "1"
    try:
        recs, n, g, exact = wl.chain_sketch(16)
        st = ez.Structure(recs, n)
    finally:
        del os.environ["EZPZ_B200_FORCE_PCG"]
    assert st.ordering()["path"] == 2
    rng = np.random.default_rng(5)
    G = g[None, :] + rng.uniform(-0.02, 0.02, (12, n))
    out = ctx.solve_batch(st, G, want_unsat=True, want_degen=True, want_jacobian=True)
    o = orc.solve_inner(recs, G[0])
    for b in range(len(G)):
        one = ctx.solve_one(st, G[b], want_jacobian=True)
        assert one.path_used == 2
        assert out.iterations[b] == one.iterations and (out.status[b] & 7) == (one.status & 7), b
        assert np.array_equal(out.final_values[b], one.final_values), b
        assert np.array_equal(out.unsat_mask[b], one.unsat_mask) and np.array_equal(out.degen_count[b], one.degen_count), b
        assert np.array_equal(out.jacobian[b], one.jacobian), b
    assert out.iterations[0] == o.iterations and np.abs(out.final_values[0] - o.final_values).max() <= 1e-9


def test_batch_of_mid_size_systems_verdicts_and_errors(ctx):
    """Inconsistent members of a batch (a contradicting Fixed row appended to the sketch) come back unsatisfied with the
    oracle's constraint list; per-problem parameter overrides are refused on this path instead of being ignored."""
    recs, n, g, exact = wl.chain_sketch(16)
    extra = ez.records([ez.Constraint.Fixed(0, float(exact[0]) + 0.5)])
    recs2 = np.concatenate([recs, extra])
    st = ez.Structure(recs2, n)
    assert st.ordering()["path"] == 1
    od = st.ordering()
    G = np.stack([g, g + 0.01])
    out = ctx.solve_batch(st, G, want_unsat=True)
    for b in range(2):
        o = orc.solve_inner_ordered(recs2, G[b], od["elim_order"], od["sum_chunk"])
        bits = np.unpackbits(out.unsat_mask[b].view(np.uint8), bitorder="little")[:st.n_cons]
        assert np.flatnonzero(bits).tolist() == o.unsatisfied and len(o.unsatisfied) > 0
        assert out.iterations[b] == o.iterations and bool(out.status[b] & 1) == o.converged and (out.status[b] & 2)
        assert_bitwise(out.final_values[b], o.final_values, f"problem {b}")
    with pytest.raises(ez.EzpzError):
        ctx.solve_batch(st, G, params=np.zeros((2, st.n_cons)))
