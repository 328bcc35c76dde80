"""tests.rs:1286-1500: the two stability / effectiveness regressions of PointArcCoincident on the fixtures
arc_line_coincident_bug (+ problem_without_arc_constraint) and arc_center_point_coincident, on the oracle and on the CUDA
path (text pipeline -> ezpz_b200_solve)."""
import math

import pytest

import orc
import textual_twin
import workloads as wl

ENGINES = [pytest.param("oracle", id="oracle"), pytest.param("gpu", id="gpu", marks=pytest.mark.gpu)]


def run(engine, name):
    """`run(name)` of tests.rs: parse, solve with analysis, look results up by label."""
    text = wl.fixture_text(name)
    if engine == "oracle":
        cs = textual_twin.parse(text)
        o = orc.solve(cs.constraints, cs.initial_guesses, analysis=True)
        assert o.rc == 0
        v = o.final_values
        pts = {lab: (v[2 * i], v[2 * i + 1]) for i, lab in enumerate(cs.inner_points)}
        a0 = 2 * len(cs.inner_points) + 3 * len(cs.inner_circles)
        arcs = {lab: {"a": (v[a0 + 6 * i], v[a0 + 6 * i + 1]), "b": (v[a0 + 6 * i + 2], v[a0 + 6 * i + 3]),
                      "center": (v[a0 + 6 * i + 4], v[a0 + 6 * i + 5])} for i, lab in enumerate(cs.inner_arcs)}
        return pts, arcs
    import ezpz_b200 as ez
    out = ez.textual.Problem(text).to_constraint_system().solve_with_config_analysis()
    return out.points, out.arcs


def dist(p, q):
    return math.hypot(p[0] - q[0], p[1] - q[1])


@pytest.mark.parametrize("engine", ENGINES)
def test_point_basically_already_on_arc_should_not_cause_much_change_in_sketch(engine):
    """tests.rs:1286-1380: adding point_arc_coincident to a point that is almost on the arc must not move it far."""
    run(engine, "arc_line_coincident_bug/problem_without_arc_constraint")  # the baseline must solve (tests.rs:1295-1303)
    pts, arcs = run(engine, "arc_line_coincident_bug")
    initial_line4_start, initial_center, initial_a = (-2.32, -2.96), (1.06, -3.26), (-1.44, -0.99)
    initial_distance_from_arc = abs(dist(initial_line4_start, initial_center) - dist(initial_center, initial_a))
    assert initial_distance_from_arc < 0.5
    change = dist(pts["line4start"], initial_line4_start)
    assert change <= initial_distance_from_arc * 10.0, change
    assert "arc1" in arcs


@pytest.mark.parametrize("engine", ENGINES)
def test_arc_center_point_coincident(engine):
    """tests.rs:1399-1500: a point outside the arc's angular range must be moved onto the arc and into the range."""
    pts, arcs = run(engine, "arc_center_point_coincident")
    p0, c0, a0, b0 = (-1.16, -2.63), (0.55, -3.31), (2.25, -3.99), (1.43, -1.71)
    start_cross0 = (a0[0] - c0[0]) * (c0[1] - p0[1]) - (a0[1] - c0[1]) * (c0[0] - p0[0])
    end_cross0 = (b0[0] - c0[0]) * (c0[1] - p0[1]) - (b0[1] - c0[1]) * (c0[0] - p0[0])
    assert not (start_cross0 <= 0.0 and end_cross0 < 0.0)
    p, arc = pts["line4start"], arcs["arc1"]
    radius = dist(arc["center"], arc["a"])
    assert abs(dist(p, arc["center"]) - radius) < 0.01
    if start_cross0 > 0.1:
        assert dist(p, p0) > radius * 0.3
    c, a, b = arc["center"], arc["a"], arc["b"]
    start_cross = (a[0] - c[0]) * (c[1] - p[1]) - (a[1] - c[1]) * (c[0] - p[0])
    end_cross = (b[0] - c[0]) * (c[1] - p[1]) - (b[1] - c[1]) * (c[0] - p[0])
    assert start_cross < 0.01 and end_cross < 1e-6
