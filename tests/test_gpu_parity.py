"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): sparsity pattern bit-exact; iteration count and converged /
unsatisfied / underconstrained verdicts identical; final coordinates within 1e-9 absolute (or relative for
large coordinates).  Residuals and Jacobian values of a single evaluation are compared BIT FOR BIT.
"""
import math

import numpy as np
import pytest

import ezpz_b200 as ez
import orc
import workloads as wl

pytestmark = pytest.mark.gpu

FIXTURES = sorted(wl.fixtures())


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def assert_bitwise(a, b, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), f"{what}: {np.count_nonzero(~same)} of {a.size} values differ, first at {np.argmin(same)}: {a.flat[np.argmin(same)]!r} vs {b.flat[np.argmin(same)]!r}"


def resolved(recs, guesses):
    """Records with Undefined sides resolved from `guesses` the way the oracle's orc_eval expects."""
    out = recs.copy()
    for r in out:
        if r["flags"] != 0:
            continue
        ids = r["ids"]
        if r["kind"] == 0:
            ux, uy = guesses[ids[2]] - guesses[ids[0]], guesses[ids[3]] - guesses[ids[1]]
            vx, vy = guesses[ids[4]] - guesses[ids[0]], guesses[ids[5]] - guesses[ids[1]]
            r["flags"] = 1 if ux * vy - uy * vx >= 0.0 else 2
        elif r["kind"] == 1:
            dist = orc.lib().orc_fn_hypot(guesses[ids[0]] - guesses[ids[3]], guesses[ids[1]] - guesses[ids[4]])
            ar, br = guesses[ids[2]], guesses[ids[5]]
            r["flags"] = 2 if abs(abs(ar - br) - dist) < abs(ar + br - dist) else 1
    return out


@pytest.mark.parametrize("name", FIXTURES)
def test_eval_bitwise_on_fixtures(ctx, name):
    """Kernel family (1)+(2): residuals and scattered Jacobian values at the fixture's initial guess."""
    recs, n, g, _ = wl.system_from_text(wl.fixture_text(name))
    st = ez.Structure(recs, n)
    r, jc, jr, dg = ctx.evaluate(st, g)
    ro, jo, dgo, pat = orc.evaluate(resolved(recs, g), n, g)
    assert_bitwise(r, ro, "residual")
    assert_bitwise(jc, jo, "jacobian (CSC)")
    # CSR order is the same values permuted
    rc, p = orc.pattern(recs, n)
    dense = {}
    for j in range(n):
        for e in range(p["csc_col_ptr"][j], p["csc_col_ptr"][j + 1]):
            dense[(p["csc_row_idx"][e], j)] = jo[e]
    k = 0
    for i in range(p["m"]):
        for e in range(p["csr_row_ptr"][i], p["csr_row_ptr"][i + 1]):
            assert bits(np.array([jr[k]]))[0] == bits(np.array([dense[(i, p["csr_col_idx"][e])]]))[0]
            k += 1
    assert np.array_equal(dg, dgo)


@pytest.mark.parametrize("name", FIXTURES)
def test_solve_fixture_matches_oracle(ctx, name):
    recs, n, g, _ = wl.system_from_text(wl.fixture_text(name))
    st = ez.Structure(recs, n)
    out = ctx.solve_one(st, g, want_jacobian=True)
    o = orc.solve_inner(recs, g, analysis=True)
    assert o.rc == 0
    assert out.iterations == o.iterations, "iteration count"
    assert out.converged == o.converged
    assert out.unsatisfied == o.unsatisfied, "inconsistent verdict"
    scale = np.maximum(1.0, np.abs(o.final_values))
    assert (np.abs(out.final_values - o.final_values) <= 1e-9 * scale).all()
    assert_bitwise(out.final_values, o.final_values, "final values")
    assert np.array_equal(out.degen_count, o.degen_count)
    mask = ctx.freedom_analysis(st, out.jacobian)[0]
    under = [j for j in range(n) if mask[j >> 5] >> (j & 31) & 1]
    assert under == o.underconstrained, "underconstrained verdict"


def test_batch_two_rectangles_matches_oracle(ctx):
    recs, n, g = wl.two_rectangles_batch(4096)
    st = ez.Structure(recs, n)
    out = ctx.solve_batch(st, g, want_degen=True)
    fin, it, status = orc.solve_batch(recs, n, g, hoist=True)
    assert np.array_equal(out.iterations, it)
    assert np.array_equal(out.status, status)
    assert_bitwise(out.final_values, fin, "final values")
    # problem 0 is the unperturbed fixture: its solution is the reference's (tests.rs:569-577)
    expect = [1, 1, 5, 1, 5, 4, 1, 4, 2, 2, 6, 2, 6, 6, 2, 6]
    assert np.abs(out.final_values[0] - expect).max() < 1e-4


def test_batch_full_size_properties(ctx):
    """Config 2 at full size (65,536): size-independent properties instead of the oracle."""
    recs, n, g = wl.two_rectangles_batch(65536)
    st = ez.Structure(recs, n)
    out = ctx.solve_batch(st, g)
    assert (out.status & 1).all(), "all converge"
    assert not (out.status & 2).any(), "all satisfied"
    expect = np.array([1, 1, 5, 1, 5, 4, 1, 4, 2, 2, 6, 2, 6, 6, 2, 6], dtype=np.float64)
    assert np.abs(out.final_values - expect[None, :]).max() < 1e-6
    # idempotence: solving from the solution takes 0 iterations and does not move
    again = ctx.solve_batch(st, out.final_values)
    assert (again.iterations == 0).all()
    assert np.array_equal(again.final_values, out.final_values)
    # a 1,024-problem sample against the oracle
    idx = np.arange(0, 65536, 64)
    fin, it, status = orc.solve_batch(recs, n, g[idx], hoist=True)
    assert np.array_equal(out.iterations[idx], it)
    assert_bitwise(out.final_values[idx], fin, "sampled final values")


@pytest.mark.parametrize("name", ["square", "circle_tangent", "arc_length", "parc_coincident", "inconsistent",
                                  "underconstrained", "perpendicular", "symmetric", "perpdist", "chamfer_square"])
def test_batch_mixed_structures(ctx, name):
    """Config 5 ingredients: perturbed batches incl. inconsistent and underconstrained systems."""
    recs, n, g = wl.perturbed_batch(name, 512, 0xE2B200D5EED00000 + 7, 0.25)
    st = ez.Structure(recs, n)
    out = ctx.solve_batch(st, g, want_degen=True, want_jacobian=True)
    fin, it, status = orc.solve_batch(recs, n, g, hoist=True)
    mism = np.flatnonzero(out.iterations != it)
    assert mism.size == 0, f"{mism.size} iteration-count mismatches, first {mism[:5]}"
    assert np.array_equal(out.status, status)
    assert_bitwise(out.final_values, fin, "final values")
    masks = ctx.freedom_analysis(st, out.jacobian)
    for b in range(0, 512, 37):
        o = orc.solve_inner(recs, g[b], analysis=True)
        under = [j for j in range(n) if masks[b, j >> 5] >> (j & 31) & 1]
        assert under == o.underconstrained


def test_params_override(ctx):
    """Per-problem targets: distance(p0,p1,d) with d varying per problem."""
    recs, n, g0, _ = wl.system_from_text(wl.fixture_text("two_rectangles"))
    B = 64
    g = np.repeat(g0[None, :], B, 0)
    params = np.repeat(recs["p0"][None, :], B, 0).copy()
    dist_rows = np.flatnonzero(recs["kind"] == ez.K_DISTANCE)
    params[:, dist_rows[0]] = np.linspace(3.0, 5.0, B)
    st = ez.Structure(recs, n)
    out = ctx.solve_batch(st, g, params=params)
    fin, it, status = orc.solve_batch(recs, n, g, params=params, hoist=True)
    assert np.array_equal(out.iterations, it)
    assert_bitwise(out.final_values, fin, "final values")
    assert np.abs((out.final_values[:, 2] - out.final_values[:, 0]) - params[:, dist_rows[0]]).max() < 1e-6


def _random_values(rng, n, scale=8.0):
    return rng.uniform(-scale, scale, n)


def test_eval_random_all_kinds_bitwise(ctx):
    """Every kind on random points (the input distribution of proptests.rs:188-234), bit for bit, plus
    degenerate inputs (coincident points) to exercise the zero-row branches."""
    rng = np.random.default_rng(20261017)
    for trial in range(40):
        cons = random_constraints(rng, 60, 24)
        recs = ez.records(cons, rng.choice([1.0, 0.5, 3.0], len(cons)))
        x = _random_values(rng, 24)
        if trial % 4 == 0:  # collapse some points onto each other
            x[2:4] = x[0:2]
            x[8:10] = x[6:8]
        st = ez.Structure(recs, 24)
        r, jc, jr, dg = ctx.evaluate(st, x)
        ro, jo, dgo, _ = orc.evaluate(resolved(recs, x), 24, x)
        assert_bitwise(r, ro, f"trial {trial} residual")
        assert_bitwise(jc, jo, f"trial {trial} jacobian")
        assert np.array_equal(dg, dgo)


def random_constraints(rng, count, n_vars):
    pts = [ez.DatumPoint.new_xy(2 * i, 2 * i + 1) for i in range(n_vars // 2)]

    def P():
        return pts[rng.integers(len(pts))]

    def Ln():
        return ez.DatumLineSegment(P(), P())

    def Ci():
        return ez.DatumCircle(P(), ez.DatumDistance(int(rng.integers(n_vars))))

    def Ar():
        return ez.DatumCircularArc(P(), P(), P())

    def AK():
        k = rng.integers(3)
        if k == 0:
            return ez.AngleKind.Parallel()
        if k == 1:
            return ez.AngleKind.Perpendicular()
        return ez.AngleKind.Other(ez.Angle.from_radians(rng.uniform(-4, 4)))

    C = ez.Constraint
    makers = [
        lambda: C.LineTangentToCircle(Ln(), Ci(), int(rng.integers(0, 3))),  # 0 = LineSide::Undefined
        lambda: C.CircleTangentToCircle(Ci(), Ci(), int(rng.integers(0, 3))),  # 0 = CircleSide::Undefined
        lambda: C.Distance(P(), P(), rng.uniform(0, 5)),
        lambda: C.DistanceVar(P(), P(), ez.DatumDistance(int(rng.integers(n_vars)))),
        lambda: C.VerticalDistance(P(), P(), rng.uniform(-5, 5)),
        lambda: C.HorizontalDistance(P(), P(), rng.uniform(-5, 5)),
        lambda: C.Vertical(Ln()),
        lambda: C.Horizontal(Ln()),
        lambda: C.LinesAtAngle(Ln(), Ln(), AK()),
        lambda: C.Fixed(int(rng.integers(n_vars)), rng.uniform(-5, 5)),
        lambda: C.ScalarEqual(int(rng.integers(n_vars)), int(rng.integers(n_vars))),
        lambda: C.PointsCoincident(P(), P()),
        lambda: C.CircleRadius(Ci(), rng.uniform(0, 5)),
        lambda: C.LinesEqualLength(Ln(), Ln()),
        lambda: C.ArcRadius(Ar(), rng.uniform(0, 5)),
        lambda: C.Arc(Ar()),
        lambda: C.Midpoint(Ln(), P()),
        lambda: C.PointLineDistance(P(), Ln(), rng.uniform(-5, 5)),
        lambda: C.VerticalPointLineDistance(P(), Ln(), rng.uniform(-5, 5)),
        lambda: C.HorizontalPointLineDistance(P(), Ln(), rng.uniform(-5, 5)),
        lambda: C.Symmetric(Ln(), P(), P()),
        lambda: C.PointArcCoincident(Ar(), P()),
        lambda: C.ArcLength(Ar(), rng.uniform(0, 9)),
        lambda: C.ArcAngle(Ar(), ez.Angle.from_degrees(rng.uniform(-200, 200))),
        lambda: C.PointsAtAngle(P(), P(), P(), AK()),
    ]
    out = [m() for m in makers]  # every kind at least once
    while len(out) < count:
        out.append(makers[rng.integers(len(makers))]())
    return out


def test_device_math_bitwise(ctx):
    """hypot / sin / cos on the device (through ArcLength, Distance evaluations) were covered above; here the
    host copies of the same functions against the oracle's."""
    rng = np.random.default_rng(7)
    L = orc.lib()
    import ctypes as C
    for _ in range(2000):
        a, b = rng.uniform(-1e3, 1e3, 2) * 10.0 ** rng.integers(-8, 8)
        assert ez.native.lib().ezpz_b200_hypot(a, b) == L.orc_fn_hypot(a, b)
        s, c = ez.angle_sincos(a)
        assert s == L.orc_fn_sin(a) and c == L.orc_fn_cos(a)


def test_api_solve_mirrors_reference_tests(ctx):
    """ezpz::solve semantics through ezpz_b200_solve (tests.rs:39-128, 256-284, 1090-1127)."""
    C, R = ez.Constraint, ez.ConstraintRequest
    # it_returns_best_satisfied_solution (tests.rs:49-68)
    reqs = [R.new(C.Fixed(0, 0.0), 0), R.new(C.Fixed(0, 1.0), 1), R.new(C.Fixed(0, 2.0), 1)]
    s = ez.solve_analysis(reqs, [(0, 0.5)])
    assert s.is_satisfied() and s.priority_solved() == 0
    # priority_solver_reports_original_indices (tests.rs:87-106)
    reqs = [R.new(C.Fixed(0, 0.0), 1), R.new(C.Fixed(0, 1.0), 0), R.new(C.Fixed(0, 2.0), 0)]
    s = ez.solve_analysis(reqs, [(0, 0.5)])
    assert s.unsatisfied() == [1, 2] and s.priority_solved() == 0
    # initials_become_finals_if_no_constraints (tests.rs:70-85)
    s = ez.solve_analysis([], [(0, 0.5)])
    assert s.is_satisfied() and list(s.final_values()) == [0.5]
    # too_many_variables / empty (tests.rs:39-47, 108-128)
    with pytest.raises(ez.FailureOutcome) as e:
        ez.solve_analysis([R.highest_priority(C.Fixed(0, 0.0))], [])
    assert e.value.name == "MissingGuess" and e.value.constraint_id == 0 and e.value.variable == 0
    # weight_biases_inconsistent_solution (tests.rs:256-284)
    reqs = [R.highest_priority(C.Fixed(0, 0.0)), R.highest_priority(C.Fixed(0, 100.0)).with_weight(100.0)]
    assert ez.solve(reqs, [(0, 50.0)]).final_values()[0] > 99.0
    reqs = [R.highest_priority(C.Fixed(0, 0.0)), R.highest_priority(C.Fixed(0, 100.0))]
    assert abs(ez.solve(reqs, [(0, 50.0)]).final_values()[0] - 50.0) < 1e-4
    # strange_nonconvergence: iterations == 2 (tests.rs:1090-1127)
    p, q, r, s_, t = [ez.DatumPoint.new_xy(2 * i, 2 * i + 1) for i in range(5)]
    reqs = [R.highest_priority(c) for c in (
        C.Fixed(0, 0.0), C.Fixed(1, 0.0), C.PointsCoincident(r, s_), C.PointsCoincident(q, p),
        C.LinesEqualLength(ez.DatumLineSegment(q, r), ez.DatumLineSegment(s_, t)))]
    g = [0.0, -0.02, -3.39, -0.38, -2.76, 4.83, -1.54, 5.21, -1.15, 2.75]
    out = ez.solve(reqs, list(enumerate(g)), ez.Config().with_max_iterations(31))
    assert out.iterations() == 2
    # warnings (tests.rs:1129-1159): lines_at_angle(..., 0rad) -> ShouldBeParallel about constraint 7
    txt = ("# constraints\npoint p\npoint q\np.x = 0\np.y = 0\nq.y = 0\nvertical(p, q)\npoint r\npoint s\nr.x = 0\n"
           "s.x = 0\ns.y = 0\nlines_at_angle(p, q, r, s, 0rad)\n\n# guesses\np roughly (3, 4)\nq roughly (5, 6)\n"
           "r roughly (3, 4)\ns roughly (5, 6)\n")
    o = ez.textual.Problem(txt).to_constraint_system().solve()
    assert any(w.about_constraint == 7 and w.content == ez.Warning.ShouldBeParallel for w in o.warnings)


def test_iteration_count_pins_on_gpu(ctx):
    """lines_at_angle_isolated / lines_angle_sign_check / points_at_angle_already_satisfied
    (tests.rs:1506-1766) through the GPU path."""
    C, R = ez.Constraint, ez.ConstraintRequest
    P = [ez.DatumPoint.new_xy(2 * i, 2 * i + 1) for i in range(4)]
    l0, l1 = ez.DatumLineSegment(P[0], P[1]), ez.DatumLineSegment(P[2], P[3])
    pi = math.pi
    cases = [([[0, 0], [1, 0], [0, 0], [0, 2]], 0.5 * pi, 0), ([[0, 0], [1, 0], [0, 0], [0, 2]], -0.5 * pi, 0),
             ([[0, 0], [1, 0], [0, 0], [2, 0]], 0.0, 0), ([[0, 0], [1, 0], [0, 0], [2, 0]], pi, 0),
             ([[0, 0], [-1, 0], [0, 0], [2, 0]], 0.0, 0), ([[0, 0], [-1, 0], [0, 0], [2, 0]], pi, 0),
             ([[0, 0], [1, 0], [0, 0], [0, 2]], 0.0, 4), ([[0, 0], [1, 0], [0, 0], [0, 2]], pi, 4),
             ([[0, 0], [0, 1], [0, 0], [0, 2]], 0.5 * pi, 4), ([[0, 0], [0, 1], [0, 0], [0, 2]], -0.5 * pi, 4)]
    cfg = ez.Config().with_max_iterations(100)
    for pts, ang, expect in cases:
        reqs = [R.highest_priority(C.LinesAtAngle(l0, l1, ez.AngleKind.Other(ez.Angle.from_radians(ang))))]
        g = [v for p in pts for v in p]
        out = ez.solve(reqs, list(enumerate(map(float, g))), cfg)
        assert out.is_satisfied() and out.iterations() == expect, (ang, out.iterations())
    l1b = ez.DatumLineSegment(P[1], P[2])
    for ang, expect in [(0.1 * pi, 3), (-0.1 * pi, 4)]:
        reqs = [R.highest_priority(c) for c in (
            C.Fixed(0, 0.0), C.Fixed(1, 0.0), C.Fixed(2, 1.0), C.Fixed(3, 0.0),
            C.LinesAtAngle(l0, l1b, ez.AngleKind.Other(ez.Angle.from_radians(ang))))]
        out = ez.solve(reqs, list(enumerate([0.0, 0.0, 1.0, 0.0, 2.0, 1.0])), cfg)
        assert out.is_satisfied() and out.iterations() == expect
    for p1, p2, ang in [([1, 0], [0, 2], 0.5 * pi), ([1, 0], [0, -2], -0.5 * pi), ([1, 0], [3, 0], 0.0),
                        ([1, 0], [-2, 0], pi), ([2, 0], [1, 1], 0.25 * pi)]:
        reqs = [R.highest_priority(C.PointsAtAngle(P[0], P[1], P[2], ez.AngleKind.Other(ez.Angle.from_radians(ang))))]
        out = ez.solve(reqs, list(enumerate(map(float, [0, 0, *p1, *p2]))), cfg)
        assert out.is_satisfied() and out.iterations() == 0


def test_batch_priority_tiers_match_oracle(ctx):
    """The priority loop of ezpz::solve (lib.rs:199-246) for a batch (ezpz_b200_solve_batch_priorities): three tiers on the
    square fixture's topology, the tier-1 target varied per problem so that some problems stop at tier 0, some at tier 1 and
    some reach tier 2; per problem identical to the CPU oracle's priority loop (values bit for bit, iterations, tier,
    unsatisfied ORIGINAL indices)."""
    recs0, n, g0, _ = wl.system_from_text(wl.fixture_text("two_rectangles"))
    recs0 = ez.records(recs0) if not isinstance(recs0, np.ndarray) else recs0
    extra = np.zeros(3, dtype=recs0.dtype)
    extra["weight"] = 1.0
    # tier 1: distance between the first rectangle's opposite corners (points 0 and 2); tier 2: two Fixed on point 4
    extra[0]["kind"], extra[0]["ids"][:4] = 2, [0, 1, 4, 5]
    extra[1]["kind"], extra[1]["ids"][0], extra[1]["p0"] = 9, 8, 2.0
    extra[2]["kind"], extra[2]["ids"][0], extra[2]["p0"] = 9, 8, 2.5   # contradicts the previous one: tier 2 never satisfied
    recs = np.concatenate([recs0, extra])
    nc = len(recs)
    prio = np.zeros(nc, np.uint32)
    prio[len(recs0)] = 1
    prio[len(recs0) + 1:] = 2
    B = 96
    rng = np.random.default_rng(7)
    g = g0[None, :] + rng.uniform(-0.2, 0.2, (B, n))
    params = np.tile(recs["p0"], (B, 1))
    base = orc.solve(recs0, g0)  # the rectangle at tier 0
    diag = float(np.hypot(base.final_values[4] - base.final_values[0], base.final_values[5] - base.final_values[1]))
    params[:, len(recs0)] = np.where(np.arange(B) % 3 == 0, diag, diag + 0.5 + 0.01 * np.arange(B))  # consistent for every third problem
    params[:, len(recs0) + 2] = np.where(np.arange(B) % 6 == 0, 2.0, 2.5)  # tier 2 consistent for every sixth
    out = ctx.solve_batch_priorities(recs, prio, n, g, params=params)
    tiers = set()
    for b in range(B):
        rb = recs.copy()
        rb["p0"] = params[b]
        o = orc.solve(rb, g[b], priorities=prio)
        got_unsat = [c for c in range(nc) if out.unsat_mask[b, c >> 5] >> (c & 31) & 1]
        assert out.iterations[b] == o.iterations and bool(out.status[b] & 1) == o.converged, b
        assert out.priority_solved[b] == o.priority_solved and got_unsat == list(o.unsatisfied), b
        assert_bitwise(out.final_values[b], o.final_values, f"problem {b}")
        tiers.add(int(out.priority_solved[b]))
    assert tiers == {0, 1, 2}


def test_solve_topology_cache(ctx):
    """ezpz_b200_solve keeps the analysed structures of the last few constraint lists (host_api.cpp): alternating
    topologies, repeated with new guesses and with a changed target (a different list -> a miss), always the oracle's answer."""
    C, R = ez.Constraint, ez.ConstraintRequest
    p, q = ez.DatumPoint.new_xy(0, 1), ez.DatumPoint.new_xy(2, 3)

    def sys_a(d):
        return [R.highest_priority(c) for c in (C.Fixed(0, 0.0), C.Fixed(1, 0.0), C.Horizontal(ez.DatumLineSegment(p, q)),
                                                C.Distance(p, q, d))]

    def sys_b():
        return [R.highest_priority(c) for c in (C.Fixed(0, 1.0), C.Fixed(1, 2.0), C.Vertical(ez.DatumLineSegment(p, q)),
                                                C.Distance(p, q, 3.0))]

    rng = np.random.default_rng(3)
    for k in range(12):
        reqs = sys_a(2.0 + (k // 6)) if k % 2 == 0 else sys_b()
        g = rng.uniform(-1.0, 4.0, 4)
        out = ez.solve(reqs, list(enumerate(g)), ctx=ctx)
        o = orc.solve(ez.records([r.constraint for r in reqs]), g)
        assert out.iterations() == o.iterations and out.is_satisfied() == (len(o.unsatisfied) == 0)
        assert_bitwise(np.array(out.final_values()), o.final_values, f"call {k}")


def test_solve_extends_a_cached_topology(ctx):
    """A constraint list that continues one ezpz_b200_solve has analysed (a constraint added to a solved sketch — the trim
    workflow of tests.rs:748-897: two arcs, then PointArcCoincident on one of them) goes through ezpz_b200_structure_extend:
    the answers are the oracle's for the longer list, call after call, as constraints are added one at a time."""
    C, R = ez.Constraint, ez.ConstraintRequest
    pts = [ez.DatumPoint.new_xy(2 * k, 2 * k + 1) for k in range(7)]
    arc1 = ez.DatumCircularArc(center=pts[0], start=pts[1], end=pts[2])
    arc2 = ez.DatumCircularArc(center=pts[3], start=pts[4], end=pts[5])
    g = np.array([30.2, 0.1, 0.3, 5.2, -0.1, -5.3, 0.2, -30.1, 5.1, 0.2, -5.2, -0.1, 0.2, 0.1])
    cons = [C.Arc(arc1), C.Arc(arc2), C.Fixed(0, 30.0), C.Fixed(1, 0.0), C.Fixed(6, 0.0), C.Fixed(7, -30.0),
            C.PointArcCoincident(arc2, pts[6]), C.PointArcCoincident(arc1, pts[6]), C.Fixed(2, 0.0)]
    for k in range(2, len(cons) + 1):
        reqs = [R.highest_priority(c) for c in cons[:k]]
        out = ez.solve(reqs, list(enumerate(g)), ctx=ctx)
        o = orc.solve(ez.records(cons[:k]), g)
        assert out.iterations() == o.iterations and out.is_satisfied() == (len(o.unsatisfied) == 0), k
        assert_bitwise(np.array(out.final_values()), o.final_values, f"{k} constraints")


def _undefined_sides_system():
    """Circles A (ids 0,1,2) and B (3,4,5), line p (6,7) - q (8,9).  A is pinned at the origin with radius 2, B has radius 1
    and its centre on the x axis, the line is horizontal between x = -5 and x = 5.  Both tangencies are given with
    Undefined sides, so each problem's guesses decide (constraints.rs:146-193): B ends up inside (|cx| = 1) or outside
    (|cx| = 3) of A, the line above (y = 2) or below (y = -2) it."""
    C = ez.Constraint
    A = ez.DatumCircle(ez.DatumPoint.new_xy(0, 1), ez.DatumDistance(2))
    B = ez.DatumCircle(ez.DatumPoint.new_xy(3, 4), ez.DatumDistance(5))
    p, q = ez.DatumPoint.new_xy(6, 7), ez.DatumPoint.new_xy(8, 9)
    ln = ez.DatumLineSegment(p, q)
    cons = [C.Fixed(0, 0.0), C.Fixed(1, 0.0), C.CircleRadius(A, 2.0), C.Fixed(4, 0.0), C.CircleRadius(B, 1.0),
            C.CircleTangentToCircle(A, B), C.Fixed(6, -5.0), C.Fixed(8, 5.0), C.Horizontal(ln), C.LineTangentToCircle(ln, A)]
    assert cons[5].flags == 0 and cons[9].flags == 0
    return ez.records(cons), 10


def test_undefined_sides_resolved_per_problem_on_device(ctx):
    """CircleSide::Undefined and LineSide::Undefined reach the device unresolved and are resolved there from each problem's
    own guesses (set_from_initial_values, constraints.rs:146-193, both branches): the batched kernel, the single solve and
    ezpz_b200_solve agree with the oracle bit for bit, and one batch holds problems of all four side combinations."""
    recs, n = _undefined_sides_system()
    st = ez.Structure(recs, n)
    rng = np.random.default_rng(146193)
    B = 512
    G = np.zeros((B, n))
    G[:, 2], G[:, 5] = 2.0 + rng.uniform(-0.2, 0.2, B), 1.0 + rng.uniform(-0.2, 0.2, B)
    G[:, 3] = rng.uniform(0.3, 4.0, B)                       # centre of B: inside or outside of A
    G[:, 6], G[:, 8] = -5.0, 5.0
    y = rng.uniform(0.5, 3.0, B) * rng.choice([-1.0, 1.0], B)  # the line: above or below
    G[:, 7], G[:, 9] = y, y + rng.uniform(-0.1, 0.1, B)
    out = ctx.solve_batch(st, G, want_unsat=True)
    fin, it, status = orc.solve_batch(recs, n, G, nthreads=2, hoist=True)
    assert np.array_equal(out.iterations, it)
    assert np.array_equal(out.status & 3, status & 3)
    assert_bitwise(out.final_values, fin, "final values")
    ok = (out.status & 3) == 1
    assert ok.mean() > 0.9
    inside = np.abs(np.abs(out.final_values[ok, 3]) - 1.0) < 1e-6
    outside = np.abs(np.abs(out.final_values[ok, 3]) - 3.0) < 1e-6
    above = np.abs(out.final_values[ok, 7] - 2.0) < 1e-6
    below = np.abs(out.final_values[ok, 7] + 2.0) < 1e-6
    assert (inside | outside).all() and (above | below).all()
    for a in (inside, outside):
        for b in (above, below):
            assert (a & b).sum() > 10, "every side combination must occur in the batch"
    # the side is decided by the sign tests of the reference, evaluated on the guesses
    dist0 = np.abs(G[ok, 3])
    want_inside = np.abs(np.abs(G[ok, 2] - G[ok, 5]) - dist0) < np.abs(G[ok, 2] + G[ok, 5] - dist0)
    assert np.array_equal(inside, want_inside)
    assert np.array_equal(below, G[ok, 7] >= 0.0) or np.array_equal(above, G[ok, 7] >= 0.0)
    for b in (0, 1, 2, 3):
        one = ctx.solve_one(st, G[b])
        o = orc.solve_inner(recs, G[b])
        assert one.iterations == o.iterations and one.converged == o.converged and one.unsatisfied == o.unsatisfied
        assert_bitwise(one.final_values, o.final_values, f"solve_one {b}")


def test_undefined_sides_mid_size_batch_and_large_path(ctx):
    """The same Undefined-side resolution on the persistent large-path kernel: 40 disjoint copies of the system above in one
    sketch (400 variables: one CTA per problem in a batch, one cluster when solved alone), every copy with its own side
    combination."""
    base, n1 = _undefined_sides_system()
    copies = 40
    recs = np.concatenate([base.copy() for _ in range(copies)])
    for k in range(copies):
        blk = recs[k * len(base):(k + 1) * len(base)]
        for c, used in enumerate([1, 1, 3, 1, 3, 6, 1, 1, 4, 7]):  # ids each kind really names; unused ids stay 0
            blk["ids"][c, :used] += n1 * k
    n = n1 * copies
    st = ez.Structure(recs, n)
    od = st.ordering()
    assert od["path"] == 1
    rng = np.random.default_rng(7)
    batch = 6
    G = np.zeros((batch, copies, n1))
    G[..., 2], G[..., 5] = 2.0, 1.0
    G[..., 3] = rng.uniform(0.3, 4.0, (batch, copies))
    G[..., 6], G[..., 8] = -5.0, 5.0
    y = rng.uniform(0.5, 3.0, (batch, copies)) * rng.choice([-1.0, 1.0], (batch, copies))
    G[..., 7], G[..., 9] = y, y
    G = G.reshape(batch, n)
    out = ctx.solve_batch(st, G, want_unsat=True)
    for b in range(batch):
        o = orc.solve_inner_ordered(recs, G[b], od["elim_order"], od["sum_chunk"])
        one = ctx.solve_one(st, G[b])
        assert out.iterations[b] == one.iterations == o.iterations
        assert bool(out.status[b] & 1) == one.converged == o.converged
        assert_bitwise(out.final_values[b], o.final_values, f"batch member {b}")
        assert_bitwise(one.final_values, o.final_values, f"single solve {b}")
    fv = out.final_values.reshape(batch, copies, n1)
    assert ((np.abs(np.abs(fv[..., 3]) - 1.0) < 1e-6).any() and (np.abs(np.abs(fv[..., 3]) - 3.0) < 1e-6).any())
