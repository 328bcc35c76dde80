"""Generates tests/golden/fixtures.json from the reference's own test inputs and known answers.

Run in the build container only (it reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_fixtures.py

What goes in:
  * the text of every test_cases/*/problem*.md (inputs of the reference's fixture-driven tests,
    ezpz/src/tests.rs:17-26), except massive_parallel_system, which tests regenerate with
    tests/workloads.py:massive_problem_text (a restatement of gen_big_problem.py);
  * the expected outcomes hand-transcribed from ezpz/src/tests.rs (file:line given per entry): point
    coordinates to 1e-4, is_satisfied, underconstrained id lists, sizes pinned by the CLI tests
    (ezpz-cli/src/main.rs:277,298).
"""
import json
import os

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

# name -> expectations (tests.rs line of the test in "src")
EXPECT = {
    "coincident": dict(src="tests.rs:131", satisfied=True, underconstrained=[], points={"p": [3, 3], "q": [3, 3]}),
    "symmetric": dict(src="tests.rs:148", satisfied=True, underconstrained=[],
                      points={"p": [0, 0], "q": [2, 2], "a": [0.5, 0.4], "b": [0.4, 0.5]}),
    "perpdist": dict(src="tests.rs:163", satisfied=True, underconstrained=[4, 5],
                     points={"p": [0, 0], "q": [2, 3], "a": [0.10055560181546289, 1.9536090405127489]}),
    "perpdist_negative": dict(src="tests.rs:189", satisfied=True, underconstrained=[4, 5],
                              points={"p": [0, 0], "q": [2, 3], "a": [1.5192717280306194, 0.476131954511605]}),
    "midpoint": dict(src="tests.rs:210", satisfied=True, underconstrained=[],
                     points={"p": [0, 0], "q": [2, 3], "m": [1, 1.5]}),
    "underconstrained": dict(src="tests.rs:221", satisfied=True, underconstrained=[0, 1],
                             points={"p": [1, 1], "q": [0, 0]}),
    "tiny": dict(src="tests.rs:233", satisfied=True, underconstrained=[], points={"p": [0, 0], "q": [0, 0]},
                 num_eqs=4, num_vars=4),
    "inconsistent": dict(src="tests.rs:242", satisfied=False, underconstrained=[],
                         points={"o": [0, 0], "p": [2.5, 2.5]}),
    "circle": dict(src="tests.rs:287", satisfied=True, underconstrained=[], points={"p": [5, 5]},
                   circles={"a": {"center": [0.1, 0.2], "radius": 3.4}}),
    "circle_center": dict(src="tests.rs:303", satisfied=True, underconstrained=[],
                          circles={"a": {"center": [0, 0], "radius": 1.0}}),
    "circle_tangent": dict(src="tests.rs:315", satisfied=True, underconstrained=[],
                           points={"p": [0, 3], "q": [5, 3]}, circle_center_y={"a": 1.5}, circle_radius={"a": 1.5}),
    "circle_tangent_other_dir": dict(src="tests.rs:329", satisfied=True, underconstrained=[],
                                     points={"p": [0, 3], "q": [5, 3]}, circle_center_y={"a": 1.5},
                                     circle_radius={"a": 1.5}),
    "two_rectangles": dict(src="tests.rs:564", satisfied=True, underconstrained=[],
                           points={"p0": [1, 1], "p1": [5, 1], "p2": [5, 4], "p3": [1, 4], "p4": [2, 2], "p5": [6, 2],
                                   "p6": [6, 6], "p7": [2, 6]}),
    "angle_parallel": dict(src="tests.rs:581", satisfied=True, underconstrained=[],
                           points={"p0": [0, 0], "p1": [4, 4], "p2": [0, 0], "p3": [4, 4]}),
    "angle_parallel_manual": dict(src="tests.rs:581", satisfied=True, underconstrained=[],
                                  points={"p0": [0, 0], "p1": [4, 4], "p2": [0, 0], "p3": [4, 4]}),
    "perpendicular": dict(src="tests.rs:594", satisfied=True, underconstrained=[],
                          points={"p0": [0, 0], "p1": [0, 4], "p2": [0, 0], "p3": [4, 0]}),
    "nonsquare": dict(src="tests.rs:605", satisfied=True, underconstrained=[], points={"p": [0, 0], "q": [0, 0]}),
    "square": dict(src="tests.rs:614", satisfied=True, underconstrained=[]),
    "parallelogram": dict(src="tests.rs:629", underconstrained=[4, 5, 6, 7]),
    "underdetermined_lines": dict(src="tests.rs:648", satisfied=True, underconstrained=[5],
                                  points={"p0": [0, 0], "p1": [4, 0], "p2": [4, 4]}),
    "arc_radius": dict(src="tests.rs:668", satisfied=True, underconstrained=[0, 1, 2, 3, 4, 5], num_eqs=4, num_vars=8),
    "parc_coincident": dict(src="tests.rs:692", satisfied=True, is_underconstrained=True),
    "arc_equidistant": dict(src="tests.rs:707", satisfied=True, underconstrained=[0, 1, 2, 3, 4, 5]),
    "chamfer_square": dict(src="tests.rs:731", satisfied=True, underconstrained=[],
                           points={"a": [0, 40], "b": [30, 40], "c": [40, 30], "d": [40, 0], "e": [0, 0]}),
    "arc_length": dict(src="tests.rs:743", satisfied=True),
    # the assertions of these three live in tests/test_reference_arc_regressions.py
    "arc_line_coincident_bug": dict(src="tests.rs:1286"),
    "arc_line_coincident_bug/problem_without_arc_constraint": dict(src="tests.rs:1295"),
    "arc_center_point_coincident": dict(src="tests.rs:1399"),
}


def main():
    out = {}
    for name, exp in EXPECT.items():
        path = os.path.join(REF, "test_cases", name + ".md") if "/" in name else os.path.join(
            REF, "test_cases", name, "problem.md")
        with open(path) as f:
            text = f.read()
        out[name] = dict(text=text, expect=exp)
    with open(os.path.join(HERE, "fixtures.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"wrote {len(out)} fixtures")


if __name__ == "__main__":
    main()
