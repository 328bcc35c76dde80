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
    """(records, n_vars, guesses, system) through the ORACLE-side reader of the text format (tests/textual_twin.py, pure
    Python, no product code): the oracle, the CPU arm of bench.py and the GPU path all receive these records, and
    test_host.py::test_text_readers_agree holds the product's own parser (textual.cpp) to the same bytes."""
    import textual_twin
    cs = textual_twin.parse(text)
    return cs.constraints, cs.num_vars, cs.initial_guesses.copy(), cs


def product_system_from_text(text):
    """The same through the product's text pipeline (ezpz_b200_problem_parse / ezpz_b200_problem_system)."""
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


def chain_sketch(cells, seed=0xE2B20004, noise=0.05):
    """Config 4 of BASELINE.json: a synthetic connected sketch of `cells` cells x 13 variables
    mixing distances, angles, circle tangents and arcs, built through the API (not text), deterministic from
    the seed.  Cell i owns a segment A_i B_i (4 vars), a circle C_i (3) and an arc R_i (6):
      A_i is chained to A_{i-1} by Horizontal/VerticalDistance and anchored by Fixed every 8th cell;
      B_i by Distance(A_i, B_i) and LinesAtAngle(line_{i-1}, line_i, theta_i) (anchor cells: Fixed B_i.y);
      C_i by CircleRadius + LineTangentToCircle(line_i, C_i) + Fixed(cx); every cell with i % 8 == 4 instead
          fixes its centre and gets its radius from CircleTangentToCircle(C_i, C_{i+1}, Exterior);
      R_i by PointsCoincident(R_i.start, B_i), ArcRadius, ArcLength, Arc and Fixed(centre.x).
    15-16 rows per 13 variables (consistent redundancy), one connected component.  `cells` must be a
    multiple of 8.  Returns (records, n_vars, guesses, exact_solution)."""
    import textual_twin as native  # (record dtype only)
    assert cells % 8 == 0 and cells >= 8
    K = cells
    i = np.arange(K)
    u = lambda k, lo, hi: lo + (uniform_pm(seed + (k << 40), K, 0.5) + 0.5) * (hi - lo)
    Ls = u(1, 2.0, 5.0)
    phi = np.deg2rad(u(2, 20.0, 70.0))
    rho = u(3, 0.5, 1.5)
    tpar = u(4, 0.2, 0.8)
    rarc = u(5, 1.0, 3.0)
    psi = np.deg2rad(u(6, 40.0, 140.0))
    alpha = u(7, 0.2, 0.8) * np.pi
    Ax, Ay = 6.0 * i, np.zeros(K)
    Bx, By = Ax + Ls * np.cos(phi), Ay + Ls * np.sin(phi)
    nx, ny = -np.sin(phi), np.cos(phi)  # left normal
    cx = Ax + tpar * (Bx - Ax) + rho * nx
    cy = Ay + tpar * (By - Ay) + rho * ny
    cr = rho.copy()
    special = (i % 8) == 4
    cx[special], cy[special] = 6.0 * i[special] + 3.0, 12.0
    nxt = (i + 1) % K
    cr[special] = np.hypot(cx[special] - cx[nxt[special]], cy[special] - cy[nxt[special]]) - cr[nxt[special]]
    sx, sy = Bx, By
    ccx, ccy = sx - rarc * np.cos(psi), sy - rarc * np.sin(psi)
    ex = ccx + np.cos(alpha) * (sx - ccx) - np.sin(alpha) * (sy - ccy)
    ey = ccy + np.sin(alpha) * (sx - ccx) + np.cos(alpha) * (sy - ccy)
    exact = np.stack([Ax, Ay, Bx, By, cx, cy, cr, sx, sy, ex, ey, ccx, ccy], axis=1).reshape(-1)
    n_vars = 13 * K
    b = 13 * i
    anchor = (i % 8) == 0
    theta = phi - np.roll(phi, 1)

    def rec(kind, ids, p0=0.0, p1=0.0, flags=0):
        count = len(ids[0]) if hasattr(ids[0], "__len__") else 1
        r = np.zeros(count, dtype=native.REC_DTYPE)
        r["kind"], r["flags"], r["p0"], r["p1"], r["weight"] = kind, flags, p0, p1, 1.0
        for k, col in enumerate(ids):
            r["ids"][:, k] = col
        return r

    per_cell = []  # list of (cell index array, records) to be merged in cell order, stable
    sel = anchor
    per_cell.append((i[sel], 0, rec(9, [b[sel]], Ax[sel])))
    per_cell.append((i[sel], 1, rec(9, [b[sel] + 1], Ay[sel])))
    sel = i > 0
    prev = b[sel] - 13
    per_cell.append((i[sel], 2, rec(5, [prev, prev + 1, b[sel], b[sel] + 1], -6.0)))
    per_cell.append((i[sel], 3, rec(4, [prev, prev + 1, b[sel], b[sel] + 1], 0.0)))
    per_cell.append((i, 4, rec(2, [b, b + 1, b + 2, b + 3], Ls)))
    sel = anchor
    per_cell.append((i[sel], 5, rec(9, [b[sel] + 3], By[sel])))
    sel = ~anchor
    prev = b[sel] - 13
    per_cell.append((i[sel], 5, rec(8, [prev, prev + 1, prev + 2, prev + 3, b[sel], b[sel] + 1, b[sel] + 2, b[sel] + 3],
                                    np.cos(theta[sel]), np.sin(theta[sel]), 2)))
    sel = ~special
    per_cell.append((i[sel], 6, rec(12, [b[sel] + 4, b[sel] + 5, b[sel] + 6], rho[sel])))
    per_cell.append((i[sel], 7, rec(0, [b[sel], b[sel] + 1, b[sel] + 2, b[sel] + 3, b[sel] + 4, b[sel] + 5, b[sel] + 6])))
    per_cell.append((i[sel], 8, rec(9, [b[sel] + 4], cx[sel])))
    sel = special
    nb = 13 * nxt[sel]
    per_cell.append((i[sel], 6, rec(9, [b[sel] + 4], cx[sel])))
    per_cell.append((i[sel], 7, rec(9, [b[sel] + 5], cy[sel])))
    per_cell.append((i[sel], 8, rec(1, [b[sel] + 4, b[sel] + 5, b[sel] + 6, nb + 4, nb + 5, nb + 6], flags=1)))
    arc = [b + 7, b + 8, b + 9, b + 10, b + 11, b + 12]
    per_cell.append((i, 9, rec(11, [b + 7, b + 8, b + 2, b + 3])))
    per_cell.append((i, 10, rec(14, arc, rarc)))
    per_cell.append((i, 11, rec(22, arc, alpha * rarc)))
    per_cell.append((i, 12, rec(15, arc)))
    per_cell.append((i, 13, rec(9, [b + 11], ccx)))
    cell_idx = np.concatenate([c for c, _, _ in per_cell])
    order_in_cell = np.concatenate([np.full(len(c), o) for c, o, _ in per_cell])
    recs = np.concatenate([r for _, _, r in per_cell])
    perm = np.lexsort((order_in_cell, cell_idx))
    recs = np.ascontiguousarray(recs[perm])
    guesses = exact + uniform_pm(seed + (99 << 40), n_vars, noise)
    return recs, n_vars, guesses, exact


def grid_truss(N, seed=0xE2B20006, noise=0.03, weights=False):
    """A 2D lattice sketch (not a chain): N x N points, Distance constraints to the right, down and one diagonal
    neighbour (a rigid triangulated truss), the first point Fixed in x and y and the second in y.  Its graph has
    separators of ~N points, so the large path's panels are hundreds of rows tall — the opposite regime of
    chain_sketch.  Returns (records, n_vars, guesses, exact_solution)."""
    import textual_twin as native  # (record dtype only)
    gx, gy = np.meshgrid(np.arange(N, dtype=np.float64) * 2.0, np.arange(N, dtype=np.float64) * 1.5, indexing="xy")
    jitter = uniform_pm(seed, 2 * N * N, 0.3).reshape(N, N, 2)
    px, py = gx + jitter[:, :, 0], gy + jitter[:, :, 1]
    exact = np.stack([px, py], axis=2).reshape(-1)
    pid = lambda r, c: 2 * (r * N + c)
    recs = []

    def rec(kind, ids, p0=0.0, w=1.0):
        r = np.zeros(1, dtype=native.REC_DTYPE)
        r["kind"], r["p0"], r["weight"] = kind, p0, w
        r["ids"][0, :len(ids)] = ids
        return r

    recs.append(rec(9, [pid(0, 0)], px[0, 0]))
    recs.append(rec(9, [pid(0, 0) + 1], py[0, 0]))
    recs.append(rec(9, [pid(0, 1) + 1], py[0, 1]))
    k = 0
    for r in range(N):
        for c in range(N):
            for dr, dc in ((0, 1), (1, 0), (1, 1)):
                r2, c2 = r + dr, c + dc
                if r2 < N and c2 < N:
                    a, b = pid(r, c), pid(r2, c2)
                    d = float(np.hypot(px[r, c] - px[r2, c2], py[r, c] - py[r2, c2]))
                    k += 1
                    recs.append(rec(2, [a, a + 1, b, b + 1], d, (0.5 if k % 3 == 0 else 2.0 if k % 3 == 1 else 1.0) if weights else 1.0))
    recs = np.ascontiguousarray(np.concatenate(recs))
    n_vars = 2 * N * N
    guesses = exact + uniform_pm(seed + (7 << 40), n_vars, noise)
    return recs, n_vars, guesses, exact
