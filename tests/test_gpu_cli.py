"""The CLI twin `ezpz_b200/_lib/ezpz-b200` run as a process, as ezpz-cli's own tests run `cargo run -- -f ...`
(ezpz-cli/src/main.rs:246-299): BASELINE.json config 1 is "test_cases/square solved via ezpz-cli"."""
import os
import re
import subprocess

import pytest

import orc
import workloads as wl

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "ezpz_b200", "_lib", "ezpz-b200")


def run_cli(tmp_path, name, *flags, stdin=False):
    text = wl.fixture_text(name)
    if stdin:
        out = subprocess.run([CLI, "-f", "-", *flags], input=text, capture_output=True, text=True, timeout=300)
    else:
        path = tmp_path / (name.replace("/", "_") + ".md")
        path.write_text(text)
        out = subprocess.run([CLI, "-f", str(path), *flags], capture_output=True, text=True, timeout=300)
    return out


def test_tiny(tmp_path):  # main.rs:259-278
    out = run_cli(tmp_path, "tiny")
    assert out.returncode == 0, out.stderr
    assert "Problem size: 4 rows, 4 vars" in out.stdout


def test_arc(tmp_path):  # main.rs:280-299
    out = run_cli(tmp_path, "arc_radius")
    assert out.returncode == 0, out.stderr
    assert "Problem size: 4 rows, 8 vars" in out.stdout


def test_tiny_inner_cases_show_points(tmp_path):  # main.rs:246-257: tiny, arc_radius, circle with show_points
    for case in ("tiny", "arc_radius", "circle"):
        out = run_cli(tmp_path, case, "--show-points")
        assert out.returncode == 0, out.stderr
        assert "Points:" in out.stdout
    assert "Circles:" in out.stdout and re.search(r"\ta: center = \(0\.10, 0\.20\), radius = 3\.40", out.stdout)


def test_square_config_1(tmp_path):
    """BASELINE.json configs[0]: iterations, verdict and points of test_cases/square, every printed line of print_output
    (main.rs:106-157) against the oracle's solve of the same file."""
    out = run_cli(tmp_path, "square", "--show-points")
    assert out.returncode == 0, out.stderr
    recs, n, g, cs = wl.system_from_text(wl.fixture_text("square"))
    o = orc.solve(recs, g)
    lines = out.stdout.splitlines()
    assert "Problem size: 10 rows, 8 vars" in lines
    assert f"Iterations needed: {o.iterations}" in lines
    assert "Solved up to priority: 0" in lines
    assert not any("did not converge" in l for l in lines) and o.converged
    assert not any("Not all constraints were satisfied" in l for l in lines) and o.unsatisfied == []
    assert any(re.fullmatch(r"Solved in \d+μs \(mean over 100 iterations\)", l) for l in lines)
    assert any(re.fullmatch(r"i\.e\. \d+ solves per second", l) for l in lines)
    at = lines.index("Points:")
    labels = ["a", "b", "c", "d"]
    for k, lab in enumerate(labels):
        x, y = o.final_values[2 * k], o.final_values[2 * k + 1]
        assert lines[at + 1 + k] == f"\t{lab}: ({x:.2f}, {y:.2f})".replace("-0.00", "0.00") or \
            lines[at + 1 + k] == f"\t{lab}: ({x:.2f}, {y:.2f})"
    # the square itself (tests.rs:614-627): a = (0,0), c = (4,4)
    assert lines[at + 1] in ("\ta: (0.00, 0.00)", "\ta: (-0.00, 0.00)", "\ta: (0.00, -0.00)", "\ta: (-0.00, -0.00)")
    assert lines[at + 3] == "\tc: (4.00, 4.00)"


def test_unsatisfied_listing_and_stdin(tmp_path):
    """print_unsatisfied (main.rs): `inconsistent` lists constraints 2..5 by index and kind; the problem may come from stdin."""
    out = run_cli(tmp_path, "inconsistent", stdin=True)
    assert out.returncode == 0, out.stderr
    assert "Not all constraints were satisfied:" in out.stdout
    listed = re.findall(r"^\t(\d+): (\w+)$", out.stdout, flags=re.M)
    assert [int(i) for i, _ in listed] == [2, 3, 4, 5]
    assert "Problem size: 6 rows, 4 vars" in out.stdout


def test_parse_error_exit_code(tmp_path):
    p = tmp_path / "bad.md"
    p.write_text("# constraints\npoint p\nnonsense(p)\n\n# guesses\np roughly (0, 0)\n")
    out = subprocess.run([CLI, "-f", str(p)], capture_output=True, text=True, timeout=60)
    assert out.returncode != 0 and "Error" in out.stderr
