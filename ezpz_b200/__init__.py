"""ezpz_b200 — Python host-side mirror of the `ezpz` crate's public API over libezpz_b200.so.

Names, argument meaning and error behaviour follow the reference (ezpz/src/lib.rs:5-18):
`Constraint`, `ConstraintRequest`, `Config`, `IdGenerator`, datums, `solve`, `solve_analysis`,
`textual.Problem` / `ConstraintSystem`.  On top of that, `Structure` and `Context` expose the batched
device API (one structure analysed once, many problems solved in one launch).

Everything numeric runs on the GPU through the C ABI; there is no CPU path in this package.
"""
import ctypes as C
import math

import numpy as np

from . import native
from .native import REC_DTYPE

EPSILON = 1e-4  # ezpz/src/lib.rs:43

# ---- kinds (ezpz/src/constraints.rs:37-93) --------------------------------------------------------
(K_LINE_TANGENT_TO_CIRCLE, K_CIRCLE_TANGENT_TO_CIRCLE, K_DISTANCE, K_DISTANCE_VAR, K_VERTICAL_DISTANCE,
 K_HORIZONTAL_DISTANCE, K_VERTICAL, K_HORIZONTAL, K_LINES_AT_ANGLE, K_FIXED, K_SCALAR_EQUAL,
 K_POINTS_COINCIDENT, K_CIRCLE_RADIUS, K_LINES_EQUAL_LENGTH, K_ARC_RADIUS, K_ARC, K_MIDPOINT,
 K_POINT_LINE_DISTANCE, K_VERTICAL_POINT_LINE_DISTANCE, K_HORIZONTAL_POINT_LINE_DISTANCE, K_SYMMETRIC,
 K_POINT_ARC_COINCIDENT, K_ARC_LENGTH, K_ARC_ANGLE, K_POINTS_AT_ANGLE) = range(25)

KIND_NAMES = ["LineTangentToCircle", "CircleTangentToCircle", "Distance", "DistanceVar", "VerticalDistance",
              "HorizontalDistance", "Vertical", "Horizontal", "LinesAtAngle", "Fixed", "ScalarEqual",
              "PointsCoincident", "CircleRadius", "LinesEqualLength", "ArcRadius", "Arc", "Midpoint",
              "PointLineDistance", "VerticalPointLineDistance", "HorizontalPointLineDistance", "Symmetric",
              "PointArcCoincident", "ArcLength", "ArcAngle", "PointsAtAngle"]
ROWS = [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 2, 1, 2, 1, 1, 1, 2, 2, 2, 1, 2]  # residual_dim


class LineSide:
    Undefined, Left, Right = 0, 1, 2


class CircleSide:
    Undefined, Exterior, Interior = 0, 1, 2


# ---- datums (ezpz/src/datatypes/inputs.rs, id.rs) -------------------------------------------------
class IdGenerator:
    def __init__(self):
        self.next = 0

    def next_id(self):
        out = self.next
        self.next += 1
        return out


class DatumDistance:
    def __init__(self, id):
        self.id = id

    def all_variables(self):
        return [self.id]


class DatumPoint:
    def __init__(self, ids=None, x_id=None, y_id=None):
        if isinstance(ids, IdGenerator):
            self.x_id, self.y_id = ids.next_id(), ids.next_id()
        else:
            self.x_id, self.y_id = x_id, y_id

    @staticmethod
    def new_xy(x, y):
        return DatumPoint(x_id=x, y_id=y)

    def id_x(self):
        return self.x_id

    def id_y(self):
        return self.y_id

    def all_variables(self):
        return [self.x_id, self.y_id]


class DatumLineSegment:
    def __init__(self, p0, p1):
        self.p0, self.p1 = p0, p1

    def all_variables(self):
        return self.p0.all_variables() + self.p1.all_variables()


class DatumCircle:
    def __init__(self, center, radius):
        self.center, self.radius = center, radius

    def all_variables(self):
        return self.center.all_variables() + [self.radius.id]


class DatumCircularArc:
    def __init__(self, center, start, end):
        self.center, self.start, self.end = center, start, end

    def all_variables(self):  # start, end, CENTER (inputs.rs:183-192)
        return self.start.all_variables() + self.end.all_variables() + self.center.all_variables()


class Angle:  # datatypes.rs:31-89
    def __init__(self, val, degrees):
        self.val, self.degrees = float(val), degrees

    @staticmethod
    def from_degrees(d):
        return Angle(d, True)

    @staticmethod
    def from_radians(r):
        return Angle(r, False)

    def to_degrees(self):
        return self.val if self.degrees else self.val * (180.0 / math.pi)

    def to_radians(self):
        return self.val * (math.pi / 180.0) if self.degrees else self.val

    def __str__(self):
        return f"{self.val}deg" if self.degrees else f"{self.val}rad"


class AngleKind:
    PARALLEL, PERPENDICULAR, OTHER = 0, 1, 2

    def __init__(self, kind, angle=None):
        self.kind, self.angle = kind, angle

    @staticmethod
    def Parallel():
        return AngleKind(AngleKind.PARALLEL)

    @staticmethod
    def Perpendicular():
        return AngleKind(AngleKind.PERPENDICULAR)

    @staticmethod
    def Other(angle):
        return AngleKind(AngleKind.OTHER, angle)


def angle_sincos(radians):
    """(sin, cos) with the bits of libm::sincos (Rotation2::from_angle_radians, vector.rs:116-120)."""
    s, c = C.c_double(), C.c_double()
    native.lib().ezpz_b200_angle_sincos(radians, C.byref(s), C.byref(c))
    return s.value, c.value


def _rot(kind):  # rotation_for_angle_kind (constraints.rs:2641-2647) -> (flags, cos, sin)
    if kind.kind == AngleKind.PARALLEL:
        return 0, 1.0, 0.0
    if kind.kind == AngleKind.PERPENDICULAR:
        return 1, 0.0, 1.0
    s, c = angle_sincos(kind.angle.to_radians())
    return 2, c, s


class Constraint:
    """One constraint in the flat 64-byte form (include/ezpz_b200.h).  Constructors are named after the
    variants of `enum Constraint` and take the same arguments in the same order."""

    def __init__(self, kind, ids, p0=0.0, p1=0.0, flags=0, angle=None):
        self.kind, self.ids, self.p0, self.p1, self.flags = kind, list(ids), float(p0), float(p1), flags
        self.angle = angle  # kept for the lint

    def residual_dim(self):
        return ROWS[self.kind]

    def constraint_kind(self):
        return KIND_NAMES[self.kind]

    def __repr__(self):
        return f"{KIND_NAMES[self.kind]}(ids={self.ids}, p0={self.p0}, p1={self.p1}, flags={self.flags})"

    # -- variants
    @staticmethod
    def LineTangentToCircle(line, circle, side=LineSide.Undefined):
        return Constraint(K_LINE_TANGENT_TO_CIRCLE, line.all_variables() + circle.all_variables(), flags=side)

    @staticmethod
    def CircleTangentToCircle(a, b, side=CircleSide.Undefined):
        return Constraint(K_CIRCLE_TANGENT_TO_CIRCLE, a.all_variables() + b.all_variables(), flags=side)

    @staticmethod
    def Distance(p0, p1, d):
        return Constraint(K_DISTANCE, p0.all_variables() + p1.all_variables(), d)

    @staticmethod
    def DistanceVar(p, q, d):
        return Constraint(K_DISTANCE_VAR, p.all_variables() + q.all_variables() + [d.id])

    @staticmethod
    def VerticalDistance(p0, p1, d):
        return Constraint(K_VERTICAL_DISTANCE, p0.all_variables() + p1.all_variables(), d)

    @staticmethod
    def HorizontalDistance(p0, p1, d):
        return Constraint(K_HORIZONTAL_DISTANCE, p0.all_variables() + p1.all_variables(), d)

    @staticmethod
    def Vertical(line):
        return Constraint(K_VERTICAL, line.all_variables())

    @staticmethod
    def Horizontal(line):
        return Constraint(K_HORIZONTAL, line.all_variables())

    @staticmethod
    def LinesAtAngle(l0, l1, kind):
        f, c, s = _rot(kind)
        return Constraint(K_LINES_AT_ANGLE, l0.all_variables() + l1.all_variables(), c, s, f, angle=kind.angle)

    @staticmethod
    def Fixed(id, v):
        return Constraint(K_FIXED, [id], v)

    @staticmethod
    def ScalarEqual(a, b):
        return Constraint(K_SCALAR_EQUAL, [a, b])

    @staticmethod
    def PointsCoincident(p0, p1):
        return Constraint(K_POINTS_COINCIDENT, p0.all_variables() + p1.all_variables())

    @staticmethod
    def CircleRadius(circle, r):
        return Constraint(K_CIRCLE_RADIUS, circle.all_variables(), r)

    @staticmethod
    def LinesEqualLength(l0, l1):
        return Constraint(K_LINES_EQUAL_LENGTH, l0.all_variables() + l1.all_variables())

    @staticmethod
    def ArcRadius(arc, r):
        return Constraint(K_ARC_RADIUS, arc.all_variables(), r)

    @staticmethod
    def Arc(arc):
        return Constraint(K_ARC, arc.all_variables())

    @staticmethod
    def Midpoint(line, point):
        return Constraint(K_MIDPOINT, line.all_variables() + point.all_variables())

    @staticmethod
    def PointLineDistance(point, line, d):
        return Constraint(K_POINT_LINE_DISTANCE, point.all_variables() + line.all_variables(), d)

    @staticmethod
    def VerticalPointLineDistance(point, line, d):
        return Constraint(K_VERTICAL_POINT_LINE_DISTANCE, point.all_variables() + line.all_variables(), d)

    @staticmethod
    def HorizontalPointLineDistance(point, line, d):
        return Constraint(K_HORIZONTAL_POINT_LINE_DISTANCE, point.all_variables() + line.all_variables(), d)

    @staticmethod
    def Symmetric(line, a, b):
        return Constraint(K_SYMMETRIC, line.all_variables() + a.all_variables() + b.all_variables())

    @staticmethod
    def PointArcCoincident(arc, point):
        return Constraint(K_POINT_ARC_COINCIDENT, arc.all_variables() + point.all_variables())

    @staticmethod
    def ArcLength(arc, d):
        return Constraint(K_ARC_LENGTH, arc.all_variables(), d)

    @staticmethod
    def ArcAngle(arc, angle):
        s, c = angle_sincos(angle.to_radians())
        return Constraint(K_ARC_ANGLE, arc.all_variables(), c, s, angle=angle)

    @staticmethod
    def PointsAtAngle(p0, p1, p2, kind):
        f, c, s = _rot(kind)
        return Constraint(K_POINTS_AT_ANGLE, p0.all_variables() + p1.all_variables() + p2.all_variables(), c, s, f,
                          angle=kind.angle)

    # -- composites (constraints/composite.rs:10-60)
    @staticmethod
    def lines_parallel(lines):
        return Constraint.LinesAtAngle(lines[0], lines[1], AngleKind.Parallel())

    @staticmethod
    def lines_perpendicular(lines):
        return Constraint.LinesAtAngle(lines[0], lines[1], AngleKind.Perpendicular())

    @staticmethod
    def point_bisects_arc(arc, point):
        return [Constraint.PointArcCoincident(arc, point),
                Constraint.Symmetric(DatumLineSegment(arc.center, point), arc.start, arc.end)]

    @staticmethod
    def parallel_lines_distance(lines, distance):
        return [Constraint.lines_parallel(lines), Constraint.PointLineDistance(lines[0].p0, lines[1], distance)]

    @staticmethod
    def circle_arc_coincident(circle, arc):
        return [Constraint.PointsCoincident(circle.center, arc.center),
                Constraint.LinesEqualLength(DatumLineSegment(arc.center, arc.start),
                                            DatumLineSegment(arc.center, arc.end))]


class ConstraintRequest:  # constraint_request.rs:13-91
    def __init__(self, constraint, priority=0, weight=1.0):
        self.constraint, self.priority, self.weight = constraint, priority, weight

    @staticmethod
    def new(constraint, priority):
        return ConstraintRequest(constraint, priority)

    @staticmethod
    def highest_priority(constraint):
        return ConstraintRequest(constraint, 0)

    def with_weight(self, weight):
        return ConstraintRequest(self.constraint, self.priority, weight)


class Config:  # solver.rs:31-81
    def __init__(self, max_iterations=35, residual_tolerance=1e-8, step_tolerance=1e-12, initial_lambda=1e-9):
        self.max_iterations, self.residual_tolerance = max_iterations, residual_tolerance
        self.step_tolerance, self.initial_lambda = step_tolerance, initial_lambda

    def _with(self, **kw):
        c = Config(self.max_iterations, self.residual_tolerance, self.step_tolerance, self.initial_lambda)
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    def with_max_iterations(self, v):
        return self._with(max_iterations=v)

    def with_convergence_tolerance(self, v):
        return self._with(residual_tolerance=v)

    def with_step_tolerance(self, v):
        return self._with(step_tolerance=v)

    def with_initial_lambda(self, v):
        return self._with(initial_lambda=v)

    def _native(self):
        return native.Config(int(self.max_iterations), float(self.residual_tolerance), float(self.step_tolerance),
                             float(self.initial_lambda))


def records(constraints, weights=None):
    """list[Constraint] -> contiguous numpy array of 64-byte records."""
    arr = np.zeros(len(constraints), dtype=REC_DTYPE)
    for i, c in enumerate(constraints):
        arr[i]["kind"] = c.kind
        arr[i]["flags"] = c.flags
        ids = list(c.ids) + [0] * (8 - len(c.ids))
        arr[i]["ids"] = ids
        arr[i]["p0"] = c.p0
        arr[i]["p1"] = c.p1
        arr[i]["weight"] = 1.0 if weights is None else weights[i]
    return arr


# ---- errors / outcomes (error.rs, solve_outcome.rs, warnings.rs) ----------------------------------
class EzpzError(Exception):
    def __init__(self, status, detail=None):
        self.status = status
        self.name = native.status_name(status)
        self.constraint_id = detail.constraint_id if detail is not None else 0
        self.variable = detail.variable if detail is not None else 0
        self.message = detail.message.decode(errors="replace") if detail is not None else ""
        super().__init__(f"{self.name}: {self.message}")


class FailureOutcome(EzpzError):
    """Err(FailureOutcome{error, warnings, num_vars, num_eqs}) (solve_outcome.rs:136-181)."""

    def __init__(self, status, detail, warnings, num_vars, num_eqs):
        super().__init__(status, detail)
        self.warnings, self.num_vars, self.num_eqs = warnings, num_vars, num_eqs


class Warning:
    Degenerate, ShouldBeParallel, ShouldBePerpendicular = 0, 1, 2

    def __init__(self, about_constraint, content, count=1, angle_deg=float("nan")):
        self.about_constraint, self.content, self.count, self.angle_deg = about_constraint, content, count, angle_deg

    def __repr__(self):
        names = ["Degenerate", "ShouldBeParallel", "ShouldBePerpendicular"]
        return f"Warning(about_constraint={self.about_constraint}, {names[self.content]})"


class SolveOutcome:
    def __init__(self, final_values, unsatisfied, iterations, converged, warnings, priority_solved,
                 underconstrained=None, path_used=0):
        self.final_values_, self.unsatisfied_ = final_values, unsatisfied
        self.iterations_, self.converged_ = iterations, converged
        self.warnings_, self.priority_solved_ = warnings, priority_solved
        self.underconstrained_ = underconstrained
        self.path_used = path_used

    def unsatisfied(self):
        return self.unsatisfied_

    def converged(self):
        return self.converged_

    def final_values(self):
        return self.final_values_

    def iterations(self):
        return self.iterations_

    def warnings(self):
        return self.warnings_

    def priority_solved(self):
        return self.priority_solved_

    def is_satisfied(self):
        return not self.unsatisfied_

    def is_unsatisfied(self):
        return bool(self.unsatisfied_)

    def final_value_point(self, p):
        return (self.final_values_[p.id_x()], self.final_values_[p.id_y()])

    def final_value_distance(self, d):
        return self.final_values_[d.id]

    def final_value_circle(self, c):
        return {"center": self.final_value_point(c.center), "radius": self.final_values_[c.radius.id]}

    def final_value_arc(self, a):
        return {"a": self.final_value_point(a.start), "b": self.final_value_point(a.end),
                "center": self.final_value_point(a.center)}

    # FreedomAnalysis (analysis.rs:22-68)
    def is_underconstrained(self):
        return bool(self.underconstrained_)

    def underconstrained(self):
        return self.underconstrained_


# ---- device objects -------------------------------------------------------------------------------
class Structure:
    """ezpz_structure_t: one sketch topology analysed once (replaces Model::new per solve)."""

    def __init__(self, recs, n_vars, var_ids=None):
        L = native.lib()
        if not isinstance(recs, np.ndarray):
            recs = records(recs)
        self.recs = np.ascontiguousarray(recs)
        self.n_cons, self.n_vars = len(self.recs), int(n_vars)
        vi = None if var_ids is None else np.ascontiguousarray(var_ids, dtype=np.uint32)
        h = C.c_void_p()
        det = native.ErrorDetail()
        rc = L.ezpz_b200_structure_create(native.ptr(self.recs), self.n_cons, native.ptr(vi), self.n_vars,
                                          C.byref(h), C.byref(det))
        if rc != 0:
            raise EzpzError(rc, det)
        self._adopt(h)

    def _adopt(self, h):
        L = native.lib()
        self.handle = h
        m, n, nj, na, nl, ncomp = C.c_uint32(), C.c_uint32(), C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint32()
        L.ezpz_b200_structure_dims(h, C.byref(m), C.byref(n), C.byref(nj), C.byref(na), None, C.byref(ncomp))
        self.m, self.n, self.nnz, self.nnz_a, self.n_components = m.value, n.value, nj.value, na.value, ncomp.value
        self._nnz_l = None

    @property
    def nnz_l(self):
        """nnz of the natural-order Cholesky factor (computed on first request for systems of the large path)."""
        if self._nnz_l is None:
            nl = C.c_uint64()
            native.lib().ezpz_b200_structure_dims(self.handle, None, None, None, None, C.byref(nl), None)
            self._nnz_l = nl.value
        return self._nnz_l

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                native.lib().ezpz_b200_structure_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def _arr(self, p, count):
        if count == 0:
            return np.zeros(0, np.uint32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(count,)).copy()

    def pattern(self):
        """The parity artefact: sorted, deduplicated pattern of J in CSC and CSR."""
        a, b, c, d = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        native.lib().ezpz_b200_structure_pattern(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        r0 = C.c_void_p()
        native.lib().ezpz_b200_structure_rows(self.handle, C.byref(r0))
        return dict(m=self.m, nnz=self.nnz, csc_col_ptr=self._arr(a, self.n + 1), csc_row_idx=self._arr(b, self.nnz),
                    csr_row_ptr=self._arr(c, self.m + 1), csr_col_idx=self._arr(d, self.nnz),
                    cons_row0=self._arr(r0, self.n_cons + 1))

    def pattern_a(self):
        a, b, c, d = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        native.lib().ezpz_b200_structure_pattern_a(self.handle, C.byref(a), C.byref(b), C.byref(c), C.byref(d))
        return dict(a_col_ptr=self._arr(a, self.n + 1), a_row_idx=self._arr(b, self.nnz_a),
                    l_col_ptr=self._arr(c, self.n + 1), l_row_idx=self._arr(d, self.nnz_l))


    def extend(self, extra):
        """The structure of this one's constraints followed by `extra` (ezpz_b200_structure_extend): a large-path system keeps
        this structure's elimination order."""
        if not isinstance(extra, np.ndarray):
            extra = records(extra)
        extra = np.ascontiguousarray(extra)
        h = C.c_void_p()
        det = native.ErrorDetail()
        rc = native.lib().ezpz_b200_structure_extend(self.handle, native.ptr(extra), len(extra), C.byref(h), C.byref(det))
        if rc != 0:
            raise EzpzError(rc, det)
        st = Structure.__new__(Structure)
        st.recs = np.concatenate([self.recs, extra])
        st.n_cons, st.n_vars = len(st.recs), self.n_vars
        st._adopt(h)
        return st

    def fingerprint(self):
        """64-bit hash of everything the host analysis produced."""
        return int(native.lib().ezpz_b200_structure_fingerprint(self.handle))

    def role_program(self, roles, stride=1):
        """The tables of the batched kernel for `roles` cooperating warps (structure.h: RoleBlob), parsed: per role its
        constraint list and its tape as a list of ops {"dst", "shape", "added", "fin", "pairs": [(a, b)...]} or {"barrier": True}
        (slots, i.e. byte offsets divided by 8 * stride)."""
        L = native.lib()
        n_words, cons_word = C.c_uint64(0), C.c_uint32(0)
        dims = (C.c_uint32 * 4)()
        rc = L.ezpz_b200_structure_role_program(self.handle, roles, stride, None, 0, C.byref(n_words), C.byref(cons_word), dims)
        if rc != 0:
            raise EzpzError(rc)
        words = np.zeros(n_words.value, dtype=np.uint32)
        L.ezpz_b200_structure_role_program(self.handle, roles, stride, native.ptr(words), n_words.value, C.byref(n_words),
                                           C.byref(cons_word), dims)
        sb = 8 * stride
        out = {"W": int(dims[0]), "n_cons": int(dims[1]), "barriers": int(dims[2]), "critical_cost": int(dims[3]), "words": words,
               "roles": []}
        for r in range(roles):
            h = words[12 * r:12 * r + 12]
            ops, w = [], int(h[2])
            for _ in range(int(h[3])):
                dst, fin, counts, last = (int(v) for v in words[w:w + 4])
                shape, n_words = last & 0xff, last >> 8
                if shape == 0:
                    ops.append({"barrier": True})
                    w += n_words
                    continue
                added, npairs = counts & 0xffff, (counts & 0xffff) + (counts >> 16)
                pairs = [(int(words[w + 4 + 2 * k]) // sb, int(words[w + 5 + 2 * k]) // sb) for k in range(npairs)]
                ops.append({"dst": dst // sb, "shape": shape, "added": added, "fin": fin // sb, "pairs": pairs})
                w += n_words
            out["roles"].append({"constraints": (words[int(h[0]):int(h[0]) + int(h[1])] & 0xffff).tolist(), "ops": ops,
                                 "x": (int(h[4]), int(h[5])), "r": (int(h[6]), int(h[7])), "j": (int(h[8]), int(h[9]))})
        return out

    def batch_shape(self, batch, sm_count=148, smem_per_block=232448):
        """(roles, problems per CTA) the batched kernel would use for `batch` problems on such a device."""
        r, t = C.c_uint32(0), C.c_uint32(0)
        rc = native.lib().ezpz_b200_structure_batch_shape(self.handle, batch, sm_count, smem_per_block, C.byref(r), C.byref(t))
        if rc != 0:
            raise EzpzError(rc)
        return int(r.value), int(t.value)

    def ordering(self):
        """How the large path solves the damped step: dict(path, elim_order, nested, n_levels, nnz_l, sum_chunk)."""
        path, nested, nl, nnz, ch, po = C.c_int32(), C.c_int32(), C.c_uint32(), C.c_uint64(), C.c_uint32(), C.c_void_p()
        native.lib().ezpz_b200_structure_ordering(self.handle, C.byref(path), C.byref(po), C.byref(nested), C.byref(nl),
                                                  C.byref(nnz), C.byref(ch))
        order = self._arr(po, self.n) if po.value else None
        return dict(path=path.value, elim_order=order, nested=bool(nested.value), n_levels=nl.value, nnz_l=nnz.value,
                    sum_chunk=ch.value)


class BatchResult:
    pass


class Context:
    """ezpz_context_t: one CUDA device + stream + workspace."""

    def __init__(self, device=0):
        h = C.c_void_p()
        det = native.ErrorDetail()
        rc = native.lib().ezpz_b200_context_create(int(device), C.byref(h), C.byref(det))
        if rc != 0:
            raise EzpzError(rc, det)
        self.handle = h
        self.device = device

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                native.lib().ezpz_b200_context_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def launches(self):
        return int(native.lib().ezpz_b200_context_launches(self.handle))

    def synchronize(self):
        native.lib().ezpz_b200_context_synchronize(self.handle)

    def solve_batch(self, st, guesses, params=None, config=None, want_unsat=True, want_degen=False, want_jacobian=False,
                    out=None, want_under=False):
        """ezpz_b200_solve_batch: host buffers in, host buffers out.  `out`: an earlier BatchResult (e.g. with
        pinned arrays) to reuse instead of allocating.  want_under: the freedom analysis fused into the call
        (BatchResult.under_mask: bit j of a problem's words = variable j underconstrained)."""
        return _batch_call(native.lib().ezpz_b200_solve_batch, self.handle, st, guesses, params, config, want_unsat, want_degen,
                           want_jacobian, out, want_under)

    def solve_batch_device(self, st, io_ptrs, batch, config=None, stream=0):
        """Device-pointer form.  io_ptrs: dict of int device addresses (guesses, final_values, iterations,
        status required)."""
        cfg = (config or Config())._native()
        io = native.BatchIO(*[C.c_void_p(io_ptrs.get(k) or None) for k in
                              ("guesses", "params", "final_values", "iterations", "status", "unsat_mask",
                               "degen_count", "jacobian", "under_mask")])
        det = native.ErrorDetail()
        rc = native.lib().ezpz_b200_solve_batch_device(self.handle, st.handle, C.byref(cfg), int(batch), C.byref(io),
                                                       C.c_void_p(stream or None), C.byref(det))
        if rc != 0:
            raise EzpzError(rc, det)

    def solve_one(self, st, guesses, config=None, want_jacobian=False):
        g = np.ascontiguousarray(guesses, dtype=np.float64)
        cfg = (config or Config())._native()
        res = BatchResult()
        res.final_values = np.empty_like(g)
        it, status, path, lin = np.zeros(1, np.uint32), np.zeros(1, np.uint8), np.zeros(1, np.int32), np.zeros(1, np.uint32)
        res.unsat_mask = np.zeros((st.n_cons + 31) // 32, np.uint32)
        res.degen_count = np.zeros(st.n_cons, np.uint32)
        res.jacobian = np.zeros(st.nnz, np.float64) if want_jacobian else None
        io = native.OneIO(native.ptr(g), native.ptr(res.final_values), native.ptr(it), native.ptr(status),
                          native.ptr(res.unsat_mask), native.ptr(res.degen_count), native.ptr(res.jacobian),
                          native.ptr(path), native.ptr(lin))
        det = native.ErrorDetail()
        rc = native.lib().ezpz_b200_solve_one(self.handle, st.handle, C.byref(cfg), C.byref(io), C.byref(det))
        if rc != 0:
            raise EzpzError(rc, det)
        res.iterations, res.status, res.path_used, res.lin_iters = int(it[0]), int(status[0]), int(path[0]), int(lin[0])
        res.converged = bool(res.status & 1)
        bits = np.unpackbits(res.unsat_mask.view(np.uint8), bitorder="little")[:st.n_cons]
        res.unsatisfied = np.flatnonzero(bits).tolist()
        return res

    def solve_batch_priorities(self, recs, priorities, n_vars, guesses, params=None, config=None):
        """The priority loop of ezpz::solve (lib.rs:199-246) for a batch of problems of one topology
        (ezpz_b200_solve_batch_priorities).  Returns a BatchResult with final_values, iterations, status,
        priority_solved and unsat_mask (ORIGINAL request indices)."""
        if not isinstance(recs, np.ndarray):
            recs = records(recs)
        recs = np.ascontiguousarray(recs)
        n_cons = len(recs)
        g = np.ascontiguousarray(guesses, dtype=np.float64).reshape(-1, n_vars)
        B = g.shape[0]
        pr = None if priorities is None else np.ascontiguousarray(priorities, dtype=np.uint32)
        pa = None if params is None else np.ascontiguousarray(params, dtype=np.float64).reshape(B, n_cons)
        cfg = (config or Config())._native()
        res = BatchResult()
        res.final_values = np.empty_like(g)
        res.iterations = np.zeros(B, np.uint32)
        res.status = np.zeros(B, np.uint8)
        res.priority_solved = np.zeros(B, np.uint32)
        res.unsat_mask = np.zeros((B, (n_cons + 31) // 32), np.uint32)
        det = native.ErrorDetail()
        rc = native.lib().ezpz_b200_solve_batch_priorities(
            self.handle, native.ptr(recs), native.ptr(pr), n_cons, int(n_vars), C.byref(cfg), B, native.ptr(g), native.ptr(pa),
            native.ptr(res.final_values), native.ptr(res.iterations), native.ptr(res.status), native.ptr(res.priority_solved),
            native.ptr(res.unsat_mask), C.byref(det))
        if rc != 0:
            raise EzpzError(rc, det)
        return res

    def time_solve_one(self, st, guesses, reps=10, final_values=None):
        """Wall-clock seconds of `reps` ezpz_b200_solve_one calls on caller-provided host buffers (pass pinned arrays
        for large systems); buffers and the ctypes structs are built once, so the time is the C call's.  Returns
        (list of seconds, iterations, status, path_used)."""
        import time
        g = np.ascontiguousarray(guesses, dtype=np.float64)
        cfg = Config()._native()
        fv = final_values if final_values is not None else np.empty_like(g)
        it, status, path, lin = np.zeros(1, np.uint32), np.zeros(1, np.uint8), np.zeros(1, np.int32), np.zeros(1, np.uint32)
        unsat = np.zeros((st.n_cons + 31) // 32, np.uint32)
        io = native.OneIO(native.ptr(g), native.ptr(fv), native.ptr(it), native.ptr(status), native.ptr(unsat), None, None,
                          native.ptr(path), native.ptr(lin))
        det = native.ErrorDetail()
        fn = native.lib().ezpz_b200_solve_one
        args = (self.handle, st.handle, C.byref(cfg), C.byref(io), C.byref(det))
        times = []
        for _ in range(reps):
            t0 = time.perf_counter()
            rc = fn(*args)
            times.append(time.perf_counter() - t0)
            if rc != 0:
                raise EzpzError(rc, det)
        return times, int(it[0]), int(status[0]), int(path[0])

    def evaluate(self, st, x):
        """One residual + Jacobian evaluation through the device assembly kernel (parity/debug entry)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        r, jc, jr = np.zeros(st.m), np.zeros(st.nnz), np.zeros(st.nnz)
        dg = np.zeros(st.n_cons, np.uint8)
        det = native.ErrorDetail()
        rc = native.lib().ezpz_b200_eval(self.handle, st.handle, native.ptr(x), native.ptr(r), native.ptr(jc),
                                         native.ptr(jr), native.ptr(dg), C.byref(det))
        if rc != 0:
            raise EzpzError(rc, det)
        return r, jc, jr, dg

    def large_bench(self, st, x, which, reps=20):
        """(mean microseconds per launch, algorithmic bytes) of a stand-alone large-path kernel:
        which 0 = assembly, 1 = SpMV J p, 2 = SpMV Jt q."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        us, by = C.c_double(), C.c_double()
        det = native.ErrorDetail()
        rc = native.lib().ezpz_b200_large_bench(self.handle, st.handle, native.ptr(x), int(which), int(reps),
                                                C.byref(us), C.byref(by), C.byref(det))
        if rc != 0:
            raise EzpzError(rc, det)
        return us.value, by.value

    def freedom_analysis(self, st, jacobian):
        j = np.ascontiguousarray(jacobian, dtype=np.float64).reshape(-1, st.nnz)
        B = j.shape[0]
        mask = np.zeros((B, (st.n_vars + 31) // 32), np.uint32)
        det = native.ErrorDetail()
        rc = native.lib().ezpz_b200_freedom_analysis(self.handle, st.handle, B, native.ptr(j), native.ptr(mask),
                                                     C.byref(det))
        if rc != 0:
            raise EzpzError(rc, det)
        return mask


def _batch_call(fn, handle, st, guesses, params, config, want_unsat, want_degen, want_jacobian, out, want_under=False):
    g = np.ascontiguousarray(guesses, dtype=np.float64).reshape(-1, st.n_vars)
    B = g.shape[0]
    cfg = (config or Config())._native()
    if out is None:
        res = BatchResult()
        res.final_values = np.empty_like(g)
        res.iterations = np.empty(B, np.uint32)
        res.status = np.empty(B, np.uint8)
        uw = (st.n_cons + 31) // 32
        res.unsat_mask = np.zeros((B, uw), np.uint32) if want_unsat else None
        res.degen_count = np.zeros((B, st.n_cons), np.uint32) if want_degen else None
        res.jacobian = np.zeros((B, st.nnz), np.float64) if want_jacobian else None
        res.under_mask = np.zeros((B, (st.n_vars + 31) // 32), np.uint32) if want_under else None
    else:
        res = out
    p = None if params is None else np.ascontiguousarray(params, dtype=np.float64)
    io = native.BatchIO(native.ptr(g), native.ptr(p), native.ptr(res.final_values), native.ptr(res.iterations),
                        native.ptr(res.status), native.ptr(res.unsat_mask), native.ptr(res.degen_count),
                        native.ptr(res.jacobian), native.ptr(getattr(res, "under_mask", None)))
    det = native.ErrorDetail()
    rc = fn(handle, st.handle, C.byref(cfg), B, C.byref(io), C.byref(det))
    if rc != 0:
        raise EzpzError(rc, det)
    return res


class MultiContext:
    """ezpz_b200_multi_*: one worker thread + context per GPU; solve_batch shards a batch over them in ONE call."""

    def __init__(self, devices=None, n_devices=0):
        h = C.c_void_p()
        det = native.ErrorDetail()
        if devices is None:
            rc = native.lib().ezpz_b200_multi_create(None, int(n_devices), C.byref(h), C.byref(det))
        else:
            arr = np.ascontiguousarray(devices, dtype=np.int32)
            rc = native.lib().ezpz_b200_multi_create(native.ptr(arr), len(arr), C.byref(h), C.byref(det))
        if rc != 0:
            raise EzpzError(rc, det)
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                native.lib().ezpz_b200_multi_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def device_count(self):
        return int(native.lib().ezpz_b200_multi_device_count(self.handle))

    @property
    def launches(self):
        return int(native.lib().ezpz_b200_multi_launches(self.handle))

    def solve_batch(self, st, guesses, params=None, config=None, want_unsat=True, want_degen=False, want_jacobian=False,
                    out=None, want_under=False):
        return _batch_call(native.lib().ezpz_b200_solve_batch_multi, self.handle, st, guesses, params, config, want_unsat,
                           want_degen, want_jacobian, out, want_under)


def _multi_solve_jobs(self, jobs, config=None, want_unsat=True, want_under=False):
    """ezpz_b200_solve_jobs_multi: several (structure, guesses[, out]) batches in ONE call — the sub-batches of a mixed
    workload.  `jobs`: list of (Structure, guesses) or (Structure, guesses, BatchResult to fill).  Returns the BatchResults."""
    cfg = (config or Config())._native()
    arr = (native.BatchJob * max(1, len(jobs)))()
    keep, results = [], []
    for k, job in enumerate(jobs):
        st, guesses = job[0], job[1]
        g = np.ascontiguousarray(guesses, dtype=np.float64).reshape(-1, st.n_vars)
        B = g.shape[0]
        res = job[2] if len(job) > 2 and job[2] is not None else None
        if res is None:
            res = BatchResult()
            res.final_values = np.empty_like(g)
            res.iterations = np.empty(B, np.uint32)
            res.status = np.empty(B, np.uint8)
            res.unsat_mask = np.zeros((B, (st.n_cons + 31) // 32), np.uint32) if want_unsat else None
            res.under_mask = np.zeros((B, (st.n_vars + 31) // 32), np.uint32) if want_under else None
            res.degen_count = res.jacobian = None
        arr[k].structure = st.handle
        arr[k].batch = B
        arr[k].io = native.BatchIO(native.ptr(g), None, native.ptr(res.final_values), native.ptr(res.iterations), native.ptr(res.status),
                                   native.ptr(res.unsat_mask), native.ptr(getattr(res, "degen_count", None)),
                                   native.ptr(getattr(res, "jacobian", None)), native.ptr(getattr(res, "under_mask", None)))
        keep.append((g, st))
        results.append(res)
    det = native.ErrorDetail()
    rc = native.lib().ezpz_b200_solve_jobs_multi(self.handle, C.byref(cfg), arr, len(jobs), C.byref(det))
    if rc != 0:
        raise EzpzError(rc, det)
    return results


MultiContext.solve_jobs = _multi_solve_jobs


class PinnedArray:
    """A numpy array in page-locked host memory every device can address (ezpz_b200_host_alloc); freed with the object."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if hasattr(shape, "__len__") else (int(shape),)
        nbytes = max(1, int(np.prod(self.shape)) * self.dtype.itemsize)
        p = C.c_void_p()
        rc = native.lib().ezpz_b200_host_alloc(nbytes, C.byref(p))
        if rc != 0:
            raise EzpzError(rc)
        self._ptr = p
        buf = (C.c_char * nbytes).from_address(p.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def __del__(self):
        try:
            if getattr(self, "_ptr", None):
                self.array = None
                native.lib().ezpz_b200_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


def pinned_batch_buffers(st, batch, want_unsat=True, want_under=False):
    """(guesses array, BatchResult) in page-locked memory for `batch` problems of `st`; keep the returned owner list alive."""
    owners = [PinnedArray((batch, st.n_vars), np.float64), PinnedArray((batch, st.n_vars), np.float64),
              PinnedArray(batch, np.uint32), PinnedArray(batch, np.uint8),
              PinnedArray((batch, (st.n_cons + 31) // 32), np.uint32)]
    res = BatchResult()
    res.final_values, res.iterations, res.status = owners[1].array, owners[2].array, owners[3].array
    res.unsat_mask = owners[4].array if want_unsat else None
    res.under_mask = None
    if want_under:
        owners.append(PinnedArray((batch, (st.n_vars + 31) // 32), np.uint32))
        res.under_mask = owners[-1].array
    res.degen_count = None
    res.jacobian = None
    return owners[0].array, res, owners


def host_register(arr):
    rc = native.lib().ezpz_b200_host_register(C.c_void_p(arr.ctypes.data), arr.nbytes)
    if rc != 0:
        raise EzpzError(rc)


def host_unregister(arr):
    native.lib().ezpz_b200_host_unregister(C.c_void_p(arr.ctypes.data))


def shard_range(batch, rank, world):
    """Contiguous shard [begin, end) of a batch for rank `rank` of `world` (ezpz_b200_shard_range)."""
    b, e = C.c_uint64(), C.c_uint64()
    native.lib().ezpz_b200_shard_range(int(batch), int(rank), int(world), C.byref(b), C.byref(e))
    return b.value, e.value


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _solve(reqs, initial_guesses, config, analysis, ctx=None):
    L = native.lib()
    cons = [r.constraint for r in reqs]
    recs = records(cons, [r.weight for r in reqs])
    n_cons = len(reqs)
    prios = np.ascontiguousarray([r.priority for r in reqs], dtype=np.uint32)
    angles = np.full(max(n_cons, 1), np.nan)
    for i, c in enumerate(cons):
        if c.kind == K_LINES_AT_ANGLE and c.flags == 2 and c.angle is not None:
            angles[i] = c.angle.to_degrees()
    ids = np.ascontiguousarray([g[0] for g in initial_guesses], dtype=np.uint32)
    vals = np.ascontiguousarray([g[1] for g in initial_guesses], dtype=np.float64)
    n_vars = len(vals)
    fv = np.zeros(max(n_vars, 1))
    un = np.zeros(max(n_cons, 1), np.uint64)
    uc = np.zeros(max(n_vars, 1), np.uint32)
    wcap = n_cons * 2 + 8
    warr = (native.WarningRec * wcap)()
    out = native.OutcomeRec()
    out.final_values, out.unsatisfied, out.underconstrained = fv.ctypes.data, un.ctypes.data, uc.ctypes.data
    out.warnings, out.warnings_cap = C.addressof(warr), wcap
    det = native.ErrorDetail()
    cfg = (config or Config())._native()
    handle = None
    if n_cons:
        handle = (ctx or default_context()).handle
    rc = L.ezpz_b200_solve(handle, native.ptr(recs) if n_cons else None, native.ptr(prios) if n_cons else None,
                           native.ptr(angles), n_cons, native.ptr(ids) if n_vars else None,
                           native.ptr(vals) if n_vars else None, n_vars, C.byref(cfg), 1 if analysis else 0,
                           C.byref(out), C.byref(det))
    warnings = [Warning(None if warr[k].about_constraint < 0 else int(warr[k].about_constraint), int(warr[k].kind),
                        int(warr[k].count), warr[k].angle_deg) for k in range(min(out.n_warnings, wcap))]
    if rc != 0:
        raise FailureOutcome(rc, det, warnings, out.num_vars, out.num_eqs)
    return SolveOutcome(fv[:n_vars].copy(), [int(v) for v in un[:out.n_unsatisfied]], int(out.iterations),
                        bool(out.converged), warnings, int(out.priority_solved),
                        [int(v) for v in uc[:out.n_underconstrained]] if analysis else None, int(out.path_used))


def solve(reqs, initial_guesses, config=None, ctx=None):
    """ezpz::solve (lib.rs:80-87).  `initial_guesses`: list of (id, value)."""
    return _solve(reqs, initial_guesses, config, False, ctx)


def solve_analysis(reqs, initial_guesses, config=None, ctx=None):
    """ezpz::solve_analysis (lib.rs:134-144)."""
    return _solve(reqs, initial_guesses, config, True, ctx)


from . import textual  # noqa: E402,F401
