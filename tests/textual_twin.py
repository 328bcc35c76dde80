"""An independent reader of the reference's text problem format for the ORACLE side of the parity tests — test infrastructure,
pure Python, no product code.

The product's parser / executor twin is C++ (ezpz_b200/csrc/textual.cpp); the tests feed the oracle through THIS reader so that
a bug in the product's label resolution or id numbering cannot hide behind "GPU == oracle" (both would otherwise receive the
same wrong records).  tests/test_host.py compares the two readers' records byte for byte on every fixture.

Follows (behaviour, not code): textual/parser.rs:29-555 (grammar), textual/executor.rs:40-445 (guess
lookup, label resolution, instruction -> constraint mapping, all priority 0 / weight 1), textual/geometry_variables.rs:92-166
(variable numbering: points 2i,2i+1; then circles cx,cy,r; then arcs a,b,center — with the reference's latent quirk that arc ids
are based at 2*num_points, ignoring circles).  Record layout: include/ezpz_b200.h (64 bytes).
"""
import math
import re

import numpy as np

REC_DTYPE = np.dtype([("kind", "<u4"), ("flags", "<u4"), ("ids", "<u4", (8,)), ("p0", "<f8"), ("p1", "<f8"), ("weight", "<f8")])

_NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|inf|nan)"
_LABEL = r"[A-Za-z0-9]+"
_CALL = re.compile(r"^([a-z_]+)\s*\((.*)\)$")
_FIX_TUPLE = re.compile(rf"^({_LABEL}(?:\.{_LABEL})?)\s*=\s*\(\s*({_NUM})\s*,\s*({_NUM})\s*\)$")
_FIX_COMP = re.compile(rf"^({_LABEL})\.([xy])\s*=\s*({_NUM})$")
_FIX_CENTER = re.compile(rf"^({_LABEL})\.center\.([xy])\s*=\s*({_NUM})$")
_GUESS_PT = re.compile(rf"^({_LABEL}(?:\.{_LABEL})?)\s+roughly\s+\(\s*({_NUM})\s*,\s*({_NUM})\s*\)$")
_GUESS_SC = re.compile(rf"^({_LABEL}(?:\.{_LABEL})?)\s+roughly\s+({_NUM})$")


class TwinError(Exception):
    pass


def _numexpr(s):
    """numexpr := num | "sqrt(" numexpr ")" (parser.rs:536-555)."""
    s = s.strip()
    if s.startswith("sqrt(") and s.endswith(")"):
        return math.sqrt(_numexpr(s[5:-1]))
    return float(s)


def _split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


class System:
    """What Problem::to_constraint_system produces: records + guesses in id order (+ the labels, for result lookup)."""

    def __init__(self, recs, guesses, points, circles, arcs, angles_deg):
        self.constraints, self.initial_guesses = recs, guesses
        self.inner_points, self.inner_circles, self.inner_arcs = points, circles, arcs
        self.angles_deg = angles_deg

    @property
    def num_vars(self):
        return len(self.initial_guesses)


def parse(text, sincos=None):
    """text -> System.  `sincos(radians) -> (sin, cos)` supplies libm's bits for lines_at_angle (default: the oracle's port)."""
    if sincos is None:
        import orc
        L = orc.lib()
        sincos = lambda r: (L.orc_fn_sin(r), L.orc_fn_cos(r))
    head, sep, tail = text.partition("# guesses")
    if not sep or "# constraints" not in head:
        raise TwinError("missing section header")
    instr_lines = [l.strip() for l in head.split("# constraints", 1)[1].splitlines() if l.strip()]
    guess_lines = [l.strip() for l in tail.splitlines() if l.strip()]

    points, circles, arcs, instrs = [], [], [], []
    for line in instr_lines:
        m = re.match(rf"^(point|circle|arc)\s+({_LABEL})$", line)
        if m:
            {"point": points, "circle": circles, "arc": arcs}[m.group(1)].append(m.group(2))
            continue
        instrs.append(line)

    pt_guess, sc_guess = {}, {}
    for line in guess_lines:
        m = _GUESS_PT.match(line)
        if m:
            pt_guess[m.group(1)] = (float(m.group(2)), float(m.group(3)))
            continue
        m = _GUESS_SC.match(line)
        if not m:
            raise TwinError(f"bad guess line: {line!r}")
        sc_guess[m.group(1)] = float(m.group(2))

    def take(d, label):
        if label not in d:
            raise TwinError(f"MissingGuess {label}")
        return d.pop(label)

    # variable numbering and guesses (executor.rs:41-108)
    g = []
    for p in points:
        g += take(pt_guess, p)
    circle_base = len(g)
    for c in circles:
        g += take(pt_guess, c + ".center")
        g.append(take(sc_guess, c + ".radius"))
    for a in arcs:
        ctr = take(pt_guess, a + ".center")
        g += take(pt_guess, a + ".a")
        g += take(pt_guess, a + ".b")
        g += ctr
    if pt_guess or sc_guess:
        raise TwinError(f"UnusedGuesses {sorted(pt_guess) + sorted(sc_guess)}")
    arc_base = 2 * len(points)  # geometry_variables.rs:92 (circles not counted: the reference's quirk, latent in the fixtures)

    def point(label):  # datum_point_for_label (executor.rs:121-174): first match wins
        if label in points:
            i = points.index(label)
            return [2 * i, 2 * i + 1]
        for i, c in enumerate(circles):
            if label == c + ".center":
                return [circle_base + 3 * i, circle_base + 3 * i + 1]
        for suffix, off in ((".center", 4), (".a", 0), (".b", 2)):
            for i, a in enumerate(arcs):
                if label == a + suffix:
                    return [arc_base + 6 * i + off, arc_base + 6 * i + off + 1]
        raise TwinError(f"UndefinedPoint {label}")

    def circle(label):  # centre then radius (executor.rs:195-206)
        c = point(label + ".center")
        if label not in circles:
            raise TwinError(f"UndefinedPoint {label}.radius")
        return c + [circle_base + 3 * circles.index(label) + 2]

    def arc(label):  # executor looks up centre, a, b; record order is start, end, centre (inputs.rs:183-192)
        c, s, e = point(label + ".center"), point(label + ".a"), point(label + ".b")
        return s + e + c

    recs, angles = [], []

    def emit(kind, ids, p0=0.0, p1=0.0, flags=0, angle=float("nan")):
        recs.append((kind, flags, list(ids) + [0] * (8 - len(ids)), p0, p1, 1.0))
        angles.append(angle)

    def fix_component(label, comp, value):  # executor.rs:259-289
        off = 0 if comp == "x" else 1
        if label in points:
            emit(9, [2 * points.index(label) + off], value)
        elif label.endswith(".center"):
            owner = label[:-len(".center")]
            if owner in circles:  # an arc's centre written this way is silently ignored by the reference
                emit(9, [circle_base + 3 * circles.index(owner) + off], value)
        else:
            raise TwinError(f"UndefinedPoint {label}")

    for line in instrs:
        m = _FIX_CENTER.match(line)
        if m:  # executor.rs:290-320
            obj, comp, val = m.group(1), m.group(2), float(m.group(3))
            off = 0 if comp == "x" else 1
            if obj in circles:
                emit(9, [circle_base + 3 * circles.index(obj) + off], val)
            elif obj in arcs:
                emit(9, [arc_base + 6 * arcs.index(obj) + 4 + off], val)
            else:
                raise TwinError(f"UndefinedPoint {obj}")
            continue
        m = _FIX_COMP.match(line)
        if m:
            fix_component(m.group(1), m.group(2), float(m.group(3)))
            continue
        m = _FIX_TUPLE.match(line)
        if m:  # parser.rs:452-471: x then y
            fix_component(m.group(1), "x", float(m.group(2)))
            fix_component(m.group(1), "y", float(m.group(3)))
            continue
        m = _CALL.match(line)
        if not m:
            raise TwinError(f"bad instruction: {line!r}")
        name, a = m.group(1), _split_args(m.group(2))
        if name == "line":
            continue
        if name == "horizontal":
            emit(7, point(a[0]) + point(a[1]))
        elif name == "vertical":
            emit(6, point(a[0]) + point(a[1]))
        elif name == "coincident":
            emit(11, point(a[0]) + point(a[1]))
        elif name == "point_arc_coincident":
            emit(21, arc(a[1]) + point(a[0]))
        elif name == "midpoint":
            emit(16, point(a[0]) + point(a[1]) + point(a[2]))
        elif name == "symmetric":
            emit(20, point(a[0]) + point(a[1]) + point(a[2]) + point(a[3]))
        elif name == "distance":
            emit(2, point(a[0]) + point(a[1]), _numexpr(a[2]))
        elif name == "parallel":
            emit(8, sum((point(x) for x in a[:4]), []), 1.0, 0.0, 0)
        elif name == "perpendicular":
            emit(8, sum((point(x) for x in a[:4]), []), 0.0, 1.0, 1)
        elif name == "lines_equal_length":
            emit(13, sum((point(x) for x in a[:4]), []))
        elif name == "lines_at_angle":
            mm = re.match(rf"^({_NUM})(deg|rad)$", a[4])
            val = float(mm.group(1))
            deg = mm.group(2) == "deg"
            rad = val * (math.pi / 180.0) if deg else val  # f64::to_radians
            s, c = sincos(rad)
            emit(8, sum((point(x) for x in a[:4]), []), c, s, 2, val if deg else val * (180.0 / math.pi))
        elif name == "radius":
            emit(12, circle(a[0]), _numexpr(a[1]))
        elif name == "tangent":
            emit(0, point(a[0]) + point(a[1]) + circle(a[2]), flags=0)
        elif name == "arc_radius":
            emit(14, arc(a[0]), float(a[1]))
        elif name == "arc_length":
            emit(22, arc(a[0]), float(a[1]))
        elif name == "is_arc":
            emit(15, arc(a[0]))
        elif name == "point_line_distance":
            emit(17, point(a[0]) + point(a[1]) + point(a[2]), float(a[3]))
        else:
            raise TwinError(f"unknown instruction {name}")
    arr = np.zeros(len(recs), dtype=REC_DTYPE)
    for i, r in enumerate(recs):
        arr[i] = r
    return System(arr, np.array(g, dtype=np.float64), points, circles, arcs, np.array(angles))
