"""The reference's property tests (ezpz/src/tests/proptests.rs) and its saved regression seeds
(ezpz/proptest-regressions/tests/proptests.txt), re-run against BOTH implementations of the path: the CPU oracle
(`-m "not gpu"`) and the CUDA path through the public `solve` / text API (`-m gpu`).  Inputs are drawn from the reference's
own ranges with a seeded generator (proptest's RNG is not reproducible here); the shrunk counter-examples recorded in the
seed file are replayed verbatim.

  proptests.rs:164-180   dependent variable ids == flattened nonzeroes            -> test_dependent_ids_match_nonzeroes
  proptests.rs:294-330   square (text problem, integer guesses in +-10000)        -> test_square
  proptests.rs:332-358   scalar_eq                                                -> test_scalar_eq
  proptests.rs:360-436   vertical / horizontal distance                           -> test_axis_distance
  proptests.rs:438-512   vertical / horizontal point-line distance                -> test_axis_point_line_distance
  proptests.rs:514-541   point on arc                                             -> test_point_arc_coincident
  proptests.rs:543-569   arc length                                               -> test_point_arc_length
  proptests.rs:571-598   circle-circle tangent (interior / exterior)              -> test_circle_circle_tangent
  proptests.rs:600-707   DistanceVar: finite partials, FD check, symmetry         -> test_distance_var_properties
The finite-difference Jacobian check over all 25 kinds (:188-234) and the scale invariance (:244-292) live in
tests/test_oracle_pins.py and tests/test_gpu_parity.py.
"""
import math

import numpy as np
import pytest

import orc
import textual_twin
import workloads as wl

EPS = 1e-4  # ezpz/src/lib.rs:43
CASES = 120
REC = textual_twin.REC_DTYPE

# ids a variant's datums hold, by kind (include/ezpz_b200.h; constraints.rs:37-93 + datatypes/inputs.rs)
N_IDS = [7, 6, 4, 5, 4, 4, 4, 4, 8, 1, 2, 4, 3, 8, 6, 6, 6, 6, 6, 6, 8, 8, 6, 6, 6]
# positions of the record's ids the residual really depends on, where that is not all of them
# (extend_dependent_variable_ids, constraints.rs:208-281: y only, x only, x only, y only, the radius only)
DEPENDENT = {4: [1, 3], 5: [0, 2], 6: [0, 2], 7: [1, 3], 12: [2]}


def rec(kind, ids, p0=0.0, p1=0.0, flags=0):
    r = np.zeros(1, dtype=REC)
    r["kind"], r["flags"], r["p0"], r["p1"], r["weight"] = kind, flags, p0, p1, 1.0
    r["ids"][0, :len(ids)] = ids
    return r


class Solved:
    def __init__(self, values, satisfied, no_warnings, iterations):
        self.v, self.satisfied, self.no_warnings, self.iterations = values, satisfied, no_warnings, iterations


def solve_with(engine, recs, guesses):
    """`ezpz::solve(requests at the highest priority, guesses, Config::default())` on either implementation."""
    recs = np.ascontiguousarray(np.concatenate(recs))
    g = np.asarray(guesses, dtype=np.float64)
    if engine == "oracle":
        o = orc.solve(recs, g)
        assert o.rc == 0
        return Solved(o.final_values, not o.unsatisfied, not o.degen_count.any(), o.iterations)
    import ezpz_b200 as ez
    from ezpz_b200 import native
    import ctypes as C
    n_cons, n_vars = len(recs), len(g)
    fv = np.zeros(n_vars)
    un = np.zeros(n_cons, np.uint64)
    warr = (native.WarningRec * 64)()
    out = native.OutcomeRec()
    out.final_values, out.unsatisfied, out.underconstrained = fv.ctypes.data, un.ctypes.data, None
    out.warnings, out.warnings_cap = C.addressof(warr), 64
    det = native.ErrorDetail()
    cfg = ez.Config()._native()
    rc = native.lib().ezpz_b200_solve(ez.default_context().handle, native.ptr(recs), None, None, n_cons, None, native.ptr(g), n_vars,
                                      C.byref(cfg), 0, C.byref(out), C.byref(det))
    assert rc == 0, det.message
    return Solved(fv, out.n_unsatisfied == 0, out.n_warnings == 0, int(out.iterations))


ENGINES = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def nearly(a, b):
    return abs(a - b) < EPS


# ---------------------------------------------------------------------------------------------------------------------
def test_dependent_ids_match_nonzeroes():
    """proptests.rs:164-180 on the oracle's pattern AND the product's host analysis (no device needed): the columns of a
    constraint's rows are exactly the ids of its datums."""
    import ezpz_b200 as ez
    rng = np.random.default_rng(164)
    for _ in range(400):
        kind = int(rng.integers(25))
        ids = rng.integers(0, 32, N_IDS[kind])
        r = rec(kind, ids, rng.uniform(-5, 5), rng.uniform(-1, 1), int(rng.integers(0, 3)) if kind in (0, 1) else 0)
        want = sorted(set(int(ids[k]) for k in DEPENDENT.get(kind, range(N_IDS[kind]))))
        rc, pat = orc.pattern(r, 32)
        assert rc == 0 and sorted(set(pat["csr_col_idx"].tolist())) == want, (kind, ids)
        assert sorted(set(ez.Structure(r, 32).pattern()["csr_col_idx"].tolist())) == want, (kind, ids)


@pytest.mark.parametrize("engine", ENGINES)
def test_square(engine):
    rng = np.random.default_rng(294)
    for _ in range(40):
        x = rng.integers(-10000, 10000, 4)
        y = rng.integers(-10000, 10000, 4)
        text = ("# constraints\n    point a\n    point b\n    point c\n    point d\n    lines_equal_length(a, b, c, d)\n"
                "    lines_equal_length(b, c, a, d)\n    horizontal(a, b)\n    vertical(b, c)\n    parallel(a, b, c, d)\n"
                "    parallel(b, c, d, a)\n    a = (0, 0)\n    c = (4, 4)\n\n    # guesses\n"
                f"    a roughly ({x[0]}, {y[0]})\n    b roughly ({x[1]}, {y[1]})\n    c roughly ({x[2]}, {y[2]})\n"
                f"    d roughly ({x[3]}, {y[3]})\n    ")
        if engine == "oracle":
            cs = textual_twin.parse(text)
            o = orc.solve(cs.constraints, cs.initial_guesses)
            assert o.rc == 0 and not o.unsatisfied, (x, y)
        else:
            import ezpz_b200 as ez
            out = ez.textual.Problem(text).to_constraint_system().solve()
            assert out.unsatisfied == [], (x, y)


@pytest.mark.parametrize("engine", ENGINES)
def test_scalar_eq(engine):
    rng = np.random.default_rng(332)
    cases = [(0.0, 0.846792320291437)] + [tuple(rng.uniform(-10, 10, 2)) for _ in range(CASES)]  # first: regression seed 1
    for gx, gy in cases:
        s = solve_with(engine, [rec(10, [0, 1])], [gx, gy])
        assert s.satisfied and s.no_warnings and nearly(s.v[0], s.v[1])


@pytest.mark.parametrize("engine", ENGINES)
def test_axis_distance(engine):
    rng = np.random.default_rng(360)
    for _ in range(CASES):
        g = rng.uniform(-100, 100, 4)  # x0 y0 x1 y1
        d = rng.uniform(0, 100)
        s = solve_with(engine, [rec(4, [0, 1, 2, 3], d)], g)  # VerticalDistance
        assert s.satisfied and s.no_warnings and nearly(s.v[1] - s.v[3], d)
        s = solve_with(engine, [rec(5, [0, 1, 2, 3], d)], g)  # HorizontalDistance
        assert s.satisfied and s.no_warnings and nearly(s.v[0] - s.v[2], d)


@pytest.mark.parametrize("engine", ENGINES)
def test_axis_point_line_distance(engine):
    rng = np.random.default_rng(438)
    # regression seed 4 (vertical): an almost horizontal line far from the point
    seeds = [((39.74751056036584, -95.46159322882576, 0.0, -95.45694757549501), (0.0, 0.0), 0.0)]
    cases = seeds + [(tuple(rng.uniform(-100, 100, 4)), tuple(rng.uniform(-100, 100, 2)), rng.uniform(0, 100)) for _ in range(CASES)]
    for (p0x, p0y, p1x, p1y), (px, py), d in cases:
        guesses = [px, py, p0x, p0y, p1x, p1y]  # point = ids 0,1; line = ids 2..5
        fixed = [rec(9, [2], p0x), rec(9, [3], p0y), rec(9, [4], p1x), rec(9, [5], p1y)]
        if abs(p1x - p0x) > EPS:  # prop_assume (proptests.rs:448)
            s = solve_with(engine, fixed + [rec(18, [0, 1, 2, 3, 4, 5], d)], guesses)
            assert s.satisfied and s.no_warnings
            slope = (s.v[5] - s.v[3]) / (s.v[4] - s.v[2])
            assert nearly(s.v[1] - (s.v[3] + slope * (s.v[0] - s.v[2])), d)
        if math.hypot(p1x - p0x, p1y - p0y) > 1e-2 and abs(p1y - p0y) > 1e-2:  # proptests.rs:495-496
            s = solve_with(engine, fixed + [rec(19, [0, 1, 2, 3, 4, 5], d)], guesses)
            assert s.satisfied and s.no_warnings
            slope = (s.v[4] - s.v[2]) / (s.v[5] - s.v[3])
            assert nearly(s.v[0] - (s.v[2] + slope * (s.v[1] - s.v[3])), d)


def _point_arc_coincident(engine, cx, cy, radius, start_deg, width_deg):
    """test_point_arc_coincident (proptests.rs:961-1073).  ids: point 0,1; centre 2,3; start 4,5; end 6,7."""
    two_pi = 2.0 * math.pi
    a0 = math.fmod(math.radians(start_deg), two_pi)
    a0 = a0 + two_pi if a0 < 0 else a0
    width = math.radians(width_deg)
    a1 = a0 + width
    sx, sy = cx + math.cos(a0) * radius, cy + math.sin(a0) * radius
    ex, ey = cx + math.cos(a1) * radius, cy + math.sin(a1) * radius
    mid = a0 + width / 2.0
    guesses = [cx + math.cos(mid) * radius, cy + math.sin(mid) * radius, cx, cy, sx, sy, ex, ey]
    arc = [4, 5, 6, 7, 2, 3]  # record order: start, end, centre
    recs = [rec(15, arc), rec(9, [2], cx), rec(9, [3], cy), rec(9, [4], sx), rec(9, [5], sy), rec(9, [6], ex), rec(9, [7], ey),
            rec(21, arc + [0, 1])]
    s = solve_with(engine, recs, guesses)
    assert s.satisfied and s.no_warnings
    ang = math.atan2(s.v[1] - cy, s.v[0] - cx) % two_pi
    if a1 <= two_pi:
        assert ang + EPS >= a0 and ang <= a1 + EPS
    else:
        assert ang + EPS >= a0 or ang <= (a1 - two_pi) + EPS
    assert nearly(math.hypot(s.v[0] - cx, s.v[1] - cy), radius)


@pytest.mark.parametrize("engine", ENGINES)
def test_point_arc_coincident(engine):
    rng = np.random.default_rng(514)
    # regression seeds 2, 3, 5 (the first two were shrunk under an older, wider range of arc spans: 5 degrees)
    seeds = [(0.0, 0.0, 1.0, 326.0065646718824, 5.0), (0.0, 0.0, 22.73229937272911, 294.58471976001573, 5.0),
             (0.0, 6.850539916263869, 19.460231588106844, 0.0, 179.95268332677125)]
    # specific_test_point_arc_coincident_off_center (proptests.rs:1256-1270 region)
    cases = seeds + [(rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(1, 50), rng.uniform(0, 360), rng.uniform(10, 350))
                     for _ in range(CASES)]
    for c in cases:
        _point_arc_coincident(engine, *c)


@pytest.mark.parametrize("engine", ENGINES)
def test_point_arc_length(engine):
    """test_point_arc_length (proptests.rs:880-957).  ids: centre 0,1; start 2,3; end 4,5."""
    rng = np.random.default_rng(543)
    two_pi = 2.0 * math.pi
    for _ in range(CASES):
        cx, cy, radius = rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(1, 50)
        start_deg, pct = rng.uniform(0, 360), rng.uniform(0.05, 0.95)
        gx, gy = rng.uniform(-10, 10), rng.uniform(-10, 10)
        if math.hypot(gx - cx, gy - cy) <= EPS:
            continue
        want = two_pi * radius * pct
        a0 = math.radians(start_deg) % two_pi
        sx, sy = cx + math.cos(a0) * radius, cy + math.sin(a0) * radius
        arc = [2, 3, 4, 5, 0, 1]
        recs = [rec(9, [0], cx), rec(9, [1], cy), rec(9, [2], sx), rec(9, [3], sy), rec(22, arc, want)]
        s = solve_with(engine, recs, [cx, cy, sx, sy, gx, gy])
        assert s.satisfied and s.no_warnings, (cx, cy, radius, start_deg, pct, gx, gy)
        assert nearly(math.hypot(s.v[4] - cx, s.v[5] - cy), radius)
        ccw = ((math.atan2(s.v[5] - cy, s.v[4] - cx) % two_pi) - a0) % two_pi
        assert nearly(radius * ccw, want)


@pytest.mark.parametrize("engine", ENGINES)
def test_circle_circle_tangent(engine):
    """test_circle_circle_tangent (proptests.rs:1075-1133).  ids: circle a 0,1,2; circle b 3,4,5."""
    rng = np.random.default_rng(571)
    for _ in range(CASES):
        ax, ay, ar, br = rng.uniform(-50, 50), rng.uniform(-50, 50), rng.uniform(1, 50), rng.uniform(1, 50)
        off, internal, positive = rng.uniform(-0.25, 0.25), bool(rng.integers(2)), bool(rng.integers(2))
        if internal and abs(ar - br) <= 1.0:
            continue
        dist = abs(ar - br) if internal else ar + br
        bx = ax + (1.0 if positive else -1.0) * (dist + off)
        recs = [rec(9, [0], ax), rec(9, [1], ay), rec(9, [2], ar), rec(9, [4], ay), rec(9, [5], br),
                rec(1, [0, 1, 2, 3, 4, 5], flags=2 if internal else 1)]
        s = solve_with(engine, recs, [ax, ay, ar, bx, ay, br])
        assert s.satisfied and s.no_warnings
        cd = math.hypot(s.v[0] - s.v[3], s.v[1] - s.v[4])
        assert nearly(cd, abs(s.v[2] - s.v[5]) if internal else s.v[2] + s.v[5])


def _eval(engine, r, x):
    """(residual, partials by id, degenerate) of one single-row constraint at x."""
    if engine == "oracle":
        res, jac, dg, pat = orc.evaluate(r, len(x), x)
        cols = pat["csr_col_idx"]
        # single row: CSR order == ascending columns; map CSC values (one row, so CSC order == ascending columns too)
        return res[0], dict(zip(cols.tolist(), jac.tolist())), bool(dg[0] & 3)
    import ezpz_b200 as ez
    st = ez.Structure(r, len(x))
    res, jc, jr, dg = ez.default_context().evaluate(st, x)
    return res[0], dict(zip(st.pattern()["csr_col_idx"].tolist(), jr.tolist())), bool(dg[0] & 3)


@pytest.mark.parametrize("engine", ENGINES)
def test_distance_var_properties(engine):
    """proptests.rs:600-707.  ids: p 0,1; q 2,3; d 4."""
    rng = np.random.default_rng(600)
    for k in range(CASES):
        px, py, qx, qy, d = rng.uniform(-100, 100, 5)
        mode = k % 3
        if mode == 0:
            qx, qy = px, py
        elif mode == 1:
            qx, qy = px + EPS * 0.5, py - EPS * 0.5
        x = np.array([px, py, qx, qy, d])
        r = rec(3, [0, 1, 2, 3, 4])
        res, pd, degen = _eval(engine, r, x)
        assert all(math.isfinite(v) for v in pd.values())  # :600-627
        swapped = rec(3, [2, 3, 0, 1, 4])
        res2, pd2, degen2 = _eval(engine, swapped, x)
        assert abs(res - res2) <= 1e-12 and degen == degen2  # :661-707
        for var in range(5):
            assert abs(pd.get(var, 0.0) - pd2.get(var, 0.0)) <= 1e-12
        if not degen:
            assert abs(pd.get(0, 0.0) + pd.get(2, 0.0)) <= 1e-12 and abs(pd.get(1, 0.0) + pd.get(3, 0.0)) <= 1e-12
        if math.hypot(px - qx, py - qy) > 1e-2:  # :629-659 central differences
            assert not degen
            for var in range(5):
                step = 1e-6 * (1.0 + abs(x[var]))
                xp, xm = x.copy(), x.copy()
                xp[var] += step
                xm[var] -= step
                num = (_eval(engine, r, xp)[0] - _eval(engine, r, xm)[0]) / (2.0 * step)
                ana = pd[var]
                assert abs(ana - num) <= 1e-6 + 1e-4 * max(abs(ana), abs(num)), (var, ana, num)
