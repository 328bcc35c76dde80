"""ezpz::textual (ezpz/src/textual.rs, textual/executor.rs:461-613): Problem -> ConstraintSystem ->
Outcome.  Parsing and the instruction -> constraint mapping run in the C++ host code
(ezpz_b200/csrc/textual.cpp); solving goes through ezpz_b200_solve."""
import ctypes as C

import numpy as np

from . import native


class TextualError(Exception):
    def __init__(self, status, detail):
        self.status = status
        self.name = native.status_name(status)
        self.message = detail.message.decode(errors="replace")
        super().__init__(f"{self.name}: {self.message}")


class Problem:
    """`Problem::from_str` (textual.rs:43-49)."""

    def __init__(self, text):
        data = text.encode()
        h = C.c_void_p()
        det = native.ErrorDetail()
        rc = native.lib().ezpz_b200_problem_parse(data, len(data), C.byref(h), C.byref(det))
        if rc != 0:
            raise TextualError(rc, det)
        self.handle = h

    from_str = classmethod(lambda cls, text: cls(text))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                native.lib().ezpz_b200_problem_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def labels(self, kind):
        L = native.lib()
        return [L.ezpz_b200_problem_label(self.handle, kind, i).decode()
                for i in range(L.ezpz_b200_problem_count(self.handle, kind))]

    def to_constraint_system(self):
        """`Problem::to_constraint_system` (executor.rs:40-445)."""
        L = native.lib()
        cons, guesses = C.c_void_p(), C.c_void_p()
        nc, nv = C.c_uint32(), C.c_uint32()
        det = native.ErrorDetail()
        rc = L.ezpz_b200_problem_system(self.handle, C.byref(cons), C.byref(nc), C.byref(guesses), C.byref(nv),
                                        C.byref(det))
        if rc != 0:
            raise TextualError(rc, det)
        recs = np.zeros(nc.value, dtype=native.REC_DTYPE)
        if nc.value:
            C.memmove(recs.ctypes.data, cons, nc.value * 64)
        g = np.zeros(nv.value)
        if nv.value:
            C.memmove(g.ctypes.data, guesses, nv.value * 8)
        ang = C.c_void_p()
        L.ezpz_b200_problem_angles_deg(self.handle, C.byref(ang))
        angles = np.full(nc.value, np.nan)
        if nc.value:
            C.memmove(angles.ctypes.data, ang, nc.value * 8)
        return ConstraintSystem(recs, g, angles, self.labels(0), self.labels(1), self.labels(2))


class Outcome:
    """`textual::Outcome` (executor.rs:588-613) (+ analysis when requested)."""

    def get_point(self, label):
        return self.points.get(label)

    def get_circle(self, label):
        return self.circles.get(label)

    def get_arc(self, label):
        return self.arcs.get(label)

    def is_satisfied(self):
        return not self.unsatisfied

    def is_unsatisfied(self):
        return bool(self.unsatisfied)


class ConstraintSystem:
    def __init__(self, recs, guesses, angles_deg, points, circles, arcs):
        self.constraints = recs  # records, all priority 0 / weight 1 (executor.rs:429-435)
        self.initial_guesses = guesses
        self.angles_deg = angles_deg
        self.inner_points, self.inner_circles, self.inner_arcs = points, circles, arcs

    @property
    def num_vars(self):
        return len(self.initial_guesses)

    @property
    def num_eqs(self):
        from . import ROWS
        return int(sum(ROWS[k] for k in self.constraints["kind"]))

    def _solve(self, config, analysis, ctx=None):
        from . import Config, FailureOutcome, Warning, default_context
        L = native.lib()
        n_cons, n_vars = len(self.constraints), len(self.initial_guesses)
        fv = np.zeros(max(n_vars, 1))
        un = np.zeros(max(n_cons, 1), np.uint64)
        uc = np.zeros(max(n_vars, 1), np.uint32)
        wcap = 2 * n_cons + 8
        warr = (native.WarningRec * wcap)()
        out = native.OutcomeRec()
        out.final_values, out.unsatisfied, out.underconstrained = fv.ctypes.data, un.ctypes.data, uc.ctypes.data
        out.warnings, out.warnings_cap = C.addressof(warr), wcap
        det = native.ErrorDetail()
        cfg = (config or Config())._native()
        handle = (ctx or default_context()).handle if n_cons else None
        rc = L.ezpz_b200_solve(handle, native.ptr(self.constraints) if n_cons else None, None,
                               native.ptr(self.angles_deg) if n_cons else None, n_cons, None,
                               native.ptr(self.initial_guesses) if n_vars else None, n_vars, C.byref(cfg),
                               1 if analysis else 0, C.byref(out), C.byref(det))
        warnings = [Warning(None if warr[k].about_constraint < 0 else int(warr[k].about_constraint),
                            int(warr[k].kind), int(warr[k].count), warr[k].angle_deg)
                    for k in range(min(out.n_warnings, wcap))]
        if rc != 0:
            raise FailureOutcome(rc, det, warnings, out.num_vars, out.num_eqs)
        o = Outcome()
        o.final_values = fv[:n_vars].copy()
        o.unsatisfied = [int(v) for v in un[:out.n_unsatisfied]]
        o.iterations = int(out.iterations)
        o.converged = bool(out.converged)
        o.priority_solved = int(out.priority_solved)
        o.warnings = warnings
        o.num_vars, o.num_eqs = int(out.num_vars), int(out.num_eqs)
        o.path_used = int(out.path_used)
        o.underconstrained = [int(v) for v in uc[:out.n_underconstrained]] if analysis else None
        # executor.rs:525-566
        o.points, o.circles, o.arcs = {}, {}, {}
        v = o.final_values
        for i, lab in enumerate(self.inner_points):
            o.points[lab] = (v[2 * i], v[2 * i + 1])
        c0 = 2 * len(self.inner_points)
        for i, lab in enumerate(self.inner_circles):
            o.circles[lab] = {"center": (v[c0 + 3 * i], v[c0 + 3 * i + 1]), "radius": v[c0 + 3 * i + 2]}
        a0 = c0 + 3 * len(self.inner_circles)
        for i, lab in enumerate(self.inner_arcs):
            b = a0 + 6 * i
            o.arcs[lab] = {"a": (v[b], v[b + 1]), "b": (v[b + 2], v[b + 3]), "center": (v[b + 4], v[b + 5])}
        return o

    def solve(self, ctx=None):
        return self._solve(None, False, ctx)

    def solve_with_config(self, config, ctx=None):
        return self._solve(config, False, ctx)

    def solve_with_config_analysis(self, config=None, ctx=None):
        return self._solve(config, True, ctx)

    def solve_no_metadata(self, config=None, ctx=None):
        return self._solve(config, False, ctx)
