"""Deterministic synthetic workloads of BASELINE.json's configs, shared by the tests,
bench.py and __graft_entry__.smoke().  Nothing here reads /root/reference."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_FIX = None

MASK64 = (1 << 64) - 1


def fixtures():
    global _FIX
    if _FIX is None:
        with open(os.path.join(HERE, "golden", "fixtures.json")) as f:
            _FIX = json.load(f)
    return _FIX


def fixture_text(name):
    return fixtures()[name]["text"]


def splitmix64(x):
    """One splitmix64 output for state value x (vectorised over numpy uint64 arrays)."""
    x = (np.asarray(x, dtype=np.uint64) + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform_pm(seed, count, half_width):
    """count doubles uniform in [-half_width, half_width): (splitmix64(seed + k) >> 11) * 2^-53 * 2hw - hw."""
    with np.errstate(over="ignore"):
        k = np.arange(count, dtype=np.uint64) + np.uint64(seed & MASK64)
        u = (splitmix64(k) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)
    return u * (2.0 * half_width) - half_width


def system_from_text(text):
    """(records, n_vars, guesses, ConstraintSystem) through the product's text pipeline."""
    import ezpz_b200 as ez
    cs = ez.textual.Problem(text).to_constraint_system()
    return cs.constraints, cs.num_vars, cs.initial_guesses.copy(), cs


def perturbed_batch(name, batch, seed, half_width=0.25):
    """`batch` copies of fixture `name` with guesses perturbed by uniform +-half_width; copy 0 unperturbed
    (config 2 of BASELINE.json: seed 0xE2B200D5EED00000 for two_rectangles)."""
    recs, n_vars, g0, _ = system_from_text(fixture_text(name))
    delta = uniform_pm(seed, batch * n_vars, half_width).reshape(batch, n_vars)
    delta[0, :] = 0.0
    return recs, n_vars, g0[None, :] + delta


def two_rectangles_batch(batch):
    return perturbed_batch("two_rectangles", batch, 0xE2B200D5EED00000)


def massive_problem_text(total_lines, overconstrain=False):
    """Restates test_cases/massive_parallel_system/gen_big_problem.py: `total_lines` independent vertical
    lines, 4 variables and 4 (or 5) rows each."""
    out = ["# constraints"]
    for line in range(total_lines):
        a, b = 2 * line, 2 * line + 1
        out += [f"point p{a}", f"point p{b}", f"vertical(p{a}, p{b})", f"p{a}.x={line}", f"p{a}.y=0", f"p{b}.y=4"]
        if overconstrain:
            out.append(f"distance(p{a}, p{b}, 4)")
    out += ["", "# guesses"]
    for line in range(total_lines):
        a, b = 2 * line, 2 * line + 1
        out += [f"p{a} roughly ({a},{a})", f"p{b} roughly ({b},{b})"]
    return "\n".join(out) + "\n"


# config 5 mix: fraction, fixture name
MIX = [(0.50, "two_rectangles"), (0.15, "square"), (0.10, "circle_tangent"), (0.05, "arc_length"),
       (0.05, "parc_coincident"), (0.05, "inconsistent"), (0.05, "underconstrained"), (0.05, "perpendicular")]


def mixed_batches(total, half_width=0.25):
    """Structure-homogeneous sub-batches of the config-5 mix: list of (name, recs, n_vars, guesses)."""
    out = []
    for idx, (frac, name) in enumerate(MIX):
        b = max(1, int(round(total * frac)))
        recs, n_vars, g = perturbed_batch(name, b, 0xE2B200D5EED00000 + (idx << 40), half_width)
        out.append((name, recs, n_vars, g))
    return out
