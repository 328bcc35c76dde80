// eval.cuh — device evaluation of one constraint: residual(s) and, when asked, the analytic partial
// derivatives, for all 25 kinds.  This is kernel family (1) of the hot path; it is shared by the
// batched small-system kernel (one thread per problem, x in shared memory) and the large-system
// assembly kernel (one thread per constraint, x in global memory) through the accessor `X`.
//
// Follows ezpz/src/constraints.rs `residual` (:499-950) and `jacobian_rows` (:1000-2293) and
// ezpz/src/vector.rs expression by expression: same operand order, same degenerate tests, no fused
// multiply-add (this translation unit MUST be compiled with -fmad=false; EZPZ_NO_FMAD is defined by
// the build together with that flag).  Residual and Jacobian are fused into one pass where the two
// reference functions compute the same sub-expressions; where they test degeneracy differently
// (e.g. Distance: the residual never is, the Jacobian is below 1e-4) both outcomes are reported.
//
// Output convention: res[] are UNWEIGHTED residuals (0 when the reference leaves them untouched);
// pd[row][k] are unweighted partials in the emission order of kinds.h; `emit[row]` is false when
// jacobian_rows emitted nothing for that row (degenerate => the row of J stays zero).
#pragma once
#ifndef EZPZ_NO_FMAD
#error "eval.cuh must be compiled with -fmad=false -DEZPZ_NO_FMAD=1 (Rust never contracts a*b+c)"
#endif

#include "dmath.cuh"
#include "kinds.h"

namespace ezd {

constexpr double kEps = 1e-4;  // lib.rs:43 EPSILON

struct V2 {
    double x, y;
};
EZ_HD V2 vsub(V2 a, V2 b) { return {a.x - b.x, a.y - b.y}; }
EZ_HD V2 vadd(V2 a, V2 b) { return {a.x + b.x, a.y + b.y}; }
EZ_HD V2 vscale(V2 a, double s) { return {a.x * s, a.y * s}; }
EZ_HD double vmag(V2 a) { return ezm::ez_hypot(a.x, a.y); }
EZ_HD double vmag2(V2 a) { return ezm::ez_pow2(a.x) + ezm::ez_pow2(a.y); }
EZ_HD double vdot(V2 a, V2 b) { return a.x * b.x + a.y * b.y; }
EZ_HD double vcross(V2 a, V2 b) { return a.x * b.y - a.y * b.x; }
// Rotation2 with col0 = (c, s): apply and apply-inverse (vector.rs:127-142)
EZ_HD V2 rot_apply(double c, double s, V2 v) { return {(c * v.x) - (s * v.y), (s * v.x) + (c * v.y)}; }
EZ_HD V2 rot_apply_inv(double c, double s, V2 v) {
    const double ns = -s;
    return {(c * v.x) - (ns * v.y), (ns * v.x) + (c * v.y)};
}

struct EvalOut {
    double res[2];
    double pd[2][8];
    bool emit[2];
    bool res_degen;
    bool jac_degen;
};

// Which part of the arc is closest (constraints.rs:2593-2606): 0 interior, 1 end, 2 start.
EZ_HD int pac_classify(V2 s, V2 e, V2 p) {
    const double a_sp = ezm::ez_wrap_0_2pi(ezm::ez_atan2(vcross(s, p), vdot(s, p)));
    const double a_se = ezm::ez_wrap_0_2pi(ezm::ez_atan2(vcross(s, e), vdot(s, e)));
    if (a_sp < a_se) return 0;
    if (vmag2(vsub(e, p)) < vmag2(vsub(s, p))) return 1;
    return 2;
}

// LinesAtAngle on four points (also ArcAngle via center->start, center->end).
template <bool JAC>
EZ_HD void lines_at_angle(double x0, double y0, double x1, double y1, double x2, double y2, double x3,
                          double y3, double c, double s, EvalOut& o) {
    const V2 u{x1 - x0, y1 - y0};
    const V2 v{x3 - x2, y3 - y2};
    const double len_u = vmag(u), len_v = vmag(v);
    if (len_u <= kEps || len_v <= kEps) {
        o.res_degen = true;
        o.jac_degen = true;
        o.emit[0] = false;
        return;
    }
    const V2 riv = rot_apply_inv(c, s, v);
    const double a = vcross(u, riv);
    const double sden = (len_u + len_v) * 0.5;
    o.res[0] = a / sden;
    if (JAC) {
        const V2 u_hat = vscale(u, 1.0 / len_u);
        const V2 v_hat = vscale(v, 1.0 / len_v);
        const double inv_s = 1.0 / sden;
        const double t = a * inv_s * 0.5;
        const V2 ru = rot_apply(c, s, u);
        const V2 df_du = vscale(vsub(V2{riv.y, -riv.x}, vscale(u_hat, t)), inv_s);
        const V2 df_dv = vscale(vsub(V2{-ru.y, ru.x}, vscale(v_hat, t)), inv_s);
        o.pd[0][0] = -df_du.x;
        o.pd[0][1] = -df_du.y;
        o.pd[0][2] = df_du.x;
        o.pd[0][3] = df_du.y;
        o.pd[0][4] = -df_dv.x;
        o.pd[0][5] = -df_dv.y;
        o.pd[0][6] = df_dv.x;
        o.pd[0][7] = df_dv.y;
    }
}

// Distance(p0, p1) residual value and its four partials into pd[row][0..4).
template <bool JAC>
EZ_HD double distance_part(double x0, double y0, double x1, double y1, double* pd, bool& emit, bool& jdeg) {
    const double dist = ezm::ez_hypot(x0 - x1, y0 - y1);
    if (JAC) {
        if (dist < kEps) {
            jdeg = true;
            emit = false;
        } else {
            pd[0] = (x0 - x1) / dist;
            pd[1] = (y0 - y1) / dist;
            pd[2] = (-x0 + x1) / dist;
            pd[3] = (-y0 + y1) / dist;
        }
    }
    return dist;
}

// `side`: resolved LineSide / CircleSide (EZPZ_SIDE_UNDEFINED behaves as Left / Exterior, exactly as the
// reference's `== Right` / `== Interior` tests do).  `X::operator()(id)` returns x[id].
template <bool JAC, class X>
EZ_HD void eval_constraint(uint32_t kind, uint32_t side, const uint32_t* __restrict__ d, double p0, double p1,
                           const X& x, EvalOut& o) {
    o.res[0] = 0.0;
    o.res[1] = 0.0;
    o.emit[0] = true;
    o.emit[1] = true;
    o.res_degen = false;
    o.jac_degen = false;
    switch (kind) {
        case EZPZ_K_LINE_TANGENT_TO_CIRCLE: {
            const V2 a{x(d[0]), x(d[1])}, b{x(d[2]), x(d[3])}, ce{x(d[4]), x(d[5])};
            const double rad = x(d[6]);
            const V2 u = vsub(b, a);
            const double mag_u = vmag(u);
            if (mag_u <= kEps) {
                o.res_degen = true;
                o.jac_degen = true;
                o.emit[0] = false;
                return;
            }
            const V2 v = vsub(ce, a);
            const double cross_uv = vcross(u, v);
            const double sgn = (side == EZPZ_LINE_SIDE_RIGHT) ? -1.0 : 1.0;
            o.res[0] = sgn * cross_uv / mag_u - ezm::ez_abs(rad);
            if (JAC) {
                const double mag3 = mag_u * mag_u * mag_u;
                const double du_x = sgn * (-(u.x * cross_uv) / mag3 + v.y / mag_u);
                const double du_y = sgn * (-(u.y * cross_uv) / mag3 - v.x / mag_u);
                const double dv_x = sgn * (-u.y / mag_u);
                const double dv_y = sgn * (u.x / mag_u);
                o.pd[0][0] = -(du_x + dv_x);
                o.pd[0][1] = -(du_y + dv_y);
                o.pd[0][2] = du_x;
                o.pd[0][3] = du_y;
                o.pd[0][4] = dv_x;
                o.pd[0][5] = dv_y;
                o.pd[0][6] = -ezm::ez_signum(rad);
            }
        } break;
        case EZPZ_K_CIRCLE_TANGENT_TO_CIRCLE: {
            const V2 ac{x(d[0]), x(d[1])}, bc{x(d[3]), x(d[4])};
            const double ar = x(d[2]), br = x(d[5]);
            const double aar = ezm::ez_abs(ar), abr = ezm::ez_abs(br);
            const double dist = vmag(vsub(ac, bc));
            o.res[0] = (side == EZPZ_CIRCLE_SIDE_INTERIOR) ? ezm::ez_abs(aar - abr) - dist : aar + abr - dist;
            if (JAC) {
                const V2 dd = vsub(bc, ac);
                const double mag_d = dist;  // |b - a| == |a - b| bit for bit (hypot takes absolute values first)
                if (mag_d <= kEps) {
                    o.jac_degen = true;
                    o.emit[0] = false;
                } else {
                    const V2 ud = vscale(dd, 1.0 / mag_d);
                    const double a_sign = ezm::ez_signum(ar), b_sign = ezm::ez_signum(br);
                    double dar, dbr;
                    if (side == EZPZ_CIRCLE_SIDE_INTERIOR) {
                        const double inner = ezm::ez_signum(aar - abr);
                        dar = inner * a_sign;
                        dbr = -inner * b_sign;
                    } else {
                        dar = a_sign;
                        dbr = b_sign;
                    }
                    o.pd[0][0] = ud.x;
                    o.pd[0][1] = ud.y;
                    o.pd[0][2] = dar;
                    o.pd[0][3] = -ud.x;
                    o.pd[0][4] = -ud.y;
                    o.pd[0][5] = dbr;
                }
            }
        } break;
        case EZPZ_K_DISTANCE: {
            const double dist = distance_part<JAC>(x(d[0]), x(d[1]), x(d[2]), x(d[3]), o.pd[0], o.emit[0], o.jac_degen);
            o.res[0] = dist - p0;
        } break;
        case EZPZ_K_DISTANCE_VAR: {
            const double px = x(d[0]), py = x(d[1]), qx = x(d[2]), qy = x(d[3]), dv = x(d[4]);
            o.res[0] = -dv + EZ_SQRT(ezm::ez_pow2(px - qx) + ezm::ez_pow2(py - qy));
            if (JAC) {
                const double dist = ezm::ez_hypot(px - qx, py - qy);
                if (dist < kEps) {
                    o.jac_degen = true;
                    o.emit[0] = false;
                } else {
                    const double rd = 1.0 / dist;
                    o.pd[0][0] = (px - qx) * rd;
                    o.pd[0][1] = (py - qy) * rd;
                    o.pd[0][2] = -(px - qx) * rd;
                    o.pd[0][3] = -(py - qy) * rd;
                    o.pd[0][4] = -1.0;
                }
            }
        } break;
        case EZPZ_K_VERTICAL_DISTANCE:
            o.res[0] = (x(d[1]) - x(d[3])) - p0;
            if (JAC) { o.pd[0][0] = 1.0; o.pd[0][1] = -1.0; }
            break;
        case EZPZ_K_HORIZONTAL_DISTANCE:
            o.res[0] = (x(d[0]) - x(d[2])) - p0;
            if (JAC) { o.pd[0][0] = 1.0; o.pd[0][1] = -1.0; }
            break;
        case EZPZ_K_VERTICAL:
            o.res[0] = x(d[0]) - x(d[2]);
            if (JAC) { o.pd[0][0] = 1.0; o.pd[0][1] = -1.0; }
            break;
        case EZPZ_K_HORIZONTAL:
            o.res[0] = x(d[1]) - x(d[3]);
            if (JAC) { o.pd[0][0] = 1.0; o.pd[0][1] = -1.0; }
            break;
        case EZPZ_K_LINES_AT_ANGLE:
            lines_at_angle<JAC>(x(d[0]), x(d[1]), x(d[2]), x(d[3]), x(d[4]), x(d[5]), x(d[6]), x(d[7]), p0, p1, o);
            break;
        case EZPZ_K_FIXED:
            o.res[0] = x(d[0]) - p0;
            if (JAC) o.pd[0][0] = 1.0;
            break;
        case EZPZ_K_SCALAR_EQUAL:
            o.res[0] = x(d[0]) - x(d[1]);
            if (JAC) { o.pd[0][0] = 1.0; o.pd[0][1] = -1.0; }
            break;
        case EZPZ_K_POINTS_COINCIDENT:
            o.res[0] = x(d[0]) - x(d[2]);
            o.res[1] = x(d[1]) - x(d[3]);
            if (JAC) {
                o.pd[0][0] = 1.0; o.pd[0][1] = -1.0;
                o.pd[1][0] = 1.0; o.pd[1][1] = -1.0;
            }
            break;
        case EZPZ_K_CIRCLE_RADIUS:
            o.res[0] = x(d[2]) - p0;
            if (JAC) o.pd[0][0] = 1.0;
            break;
        case EZPZ_K_LINES_EQUAL_LENGTH: {
            const double x0 = x(d[0]), y0 = x(d[1]), x1 = x(d[2]), y1 = x(d[3]);
            const double x2 = x(d[4]), y2 = x(d[5]), x3 = x(d[6]), y3 = x(d[7]);
            const double len0 = ezm::ez_hypot(x0 - x1, y0 - y1);
            const double len1 = ezm::ez_hypot(x2 - x3, y2 - y3);
            o.res[0] = len0 - len1;
            if (JAC) {
                if (len0 < kEps || len1 < kEps) {
                    o.jac_degen = true;
                    o.emit[0] = false;
                } else {
                    o.pd[0][0] = (x0 - x1) / len0;
                    o.pd[0][1] = (y0 - y1) / len0;
                    o.pd[0][2] = (-x0 + x1) / len0;
                    o.pd[0][3] = (-y0 + y1) / len0;
                    o.pd[0][4] = (-x2 + x3) / len1;
                    o.pd[0][5] = (-y2 + y3) / len1;
                    o.pd[0][6] = (x2 - x3) / len1;
                    o.pd[0][7] = (y2 - y3) / len1;
                }
            }
        } break;
        case EZPZ_K_ARC_RADIUS: {  // Distance(center, start) then Distance(center, end)
            const double sx = x(d[0]), sy = x(d[1]), ex = x(d[2]), ey = x(d[3]), cx = x(d[4]), cy = x(d[5]);
            o.res[0] = distance_part<JAC>(cx, cy, sx, sy, o.pd[0], o.emit[0], o.jac_degen) - p0;
            o.res[1] = distance_part<JAC>(cx, cy, ex, ey, o.pd[1], o.emit[1], o.jac_degen) - p0;
        } break;
        case EZPZ_K_ARC: {
            const double sx = x(d[0]), sy = x(d[1]), ex = x(d[2]), ey = x(d[3]), cx = x(d[4]), cy = x(d[5]);
            const double usx = sx - cx, usy = sy - cy, uex = ex - cx, uey = ey - cy;
            const double dist0 = ezm::ez_hypot(usx, usy), dist1 = ezm::ez_hypot(uex, uey);
            o.res[0] = dist0 - dist1;
            if (JAC) {
                if (dist0 <= kEps || dist1 <= kEps) {
                    o.jac_degen = true;
                    o.emit[0] = false;
                } else {
                    o.pd[0][0] = usx / dist0;
                    o.pd[0][1] = usy / dist0;
                    o.pd[0][2] = -uex / dist1;
                    o.pd[0][3] = -uey / dist1;
                    o.pd[0][4] = -usx / dist0 + uex / dist1;
                    o.pd[0][5] = -usy / dist0 + uey / dist1;
                }
            }
        } break;
        case EZPZ_K_MIDPOINT: {
            const double px = x(d[0]), py = x(d[1]), qx = x(d[2]), qy = x(d[3]), ax = x(d[4]), ay = x(d[5]);
            o.res[0] = ax - px / 2.0 - qx / 2.0;
            o.res[1] = ay - py / 2.0 - qy / 2.0;
            if (JAC) {
                o.pd[0][0] = 1.0; o.pd[0][1] = -0.5; o.pd[0][2] = -0.5;
                o.pd[1][0] = 1.0; o.pd[1][1] = -0.5; o.pd[1][2] = -0.5;
            }
        } break;
        case EZPZ_K_POINT_LINE_DISTANCE: {
            const double px = x(d[0]), py = x(d[1]);
            const double p0x = x(d[2]), p0y = x(d[3]), p1x = x(d[4]), p1y = x(d[5]);
            const double a = p0y - p1y, b = p1x - p0x, cc = (p0x * p1y) - (p1x * p0y);
            const double den = ezm::ez_hypot(a, b);
            if (den < kEps) {
                o.res_degen = true;  // residual stays 0; the Jacobian has no guard (constraints.rs:2455-2489)
            } else {
                o.res[0] = (a * px + b * py + cc) / den - p0;
            }
            if (JAC) {
                const double ed = ezm::ez_hypot(-p0x + p1x, p0y - p1y);
                const double dn = ezm::ez_pow_1p5(ezm::ez_pow2(-p0x + p1x) + ezm::ez_pow2(p0y - p1y));
                const double common = p0x * p1y - p0y * p1x + px * (p0y - p1y) + py * (-p0x + p1x);
                o.pd[0][0] = (p0y - p1y) / ed;
                o.pd[0][1] = (-p0x + p1x) / ed;
                o.pd[0][2] = ((-p0x + p1x) * common) / dn + (p1y - py) / ed;
                o.pd[0][3] = ((-p0y + p1y) * common) / dn + (-p1x + px) / ed;
                o.pd[0][4] = ((p0x - p1x) * common) / dn + (-p0y + py) / ed;
                o.pd[0][5] = ((p0y - p1y) * common) / dn + (p0x - px) / ed;
            }
        } break;
        case EZPZ_K_VERTICAL_POINT_LINE_DISTANCE: {
            const double ax = x(d[0]), ay = x(d[1]), px = x(d[2]), py = x(d[3]), qx = x(d[4]), qy = x(d[5]);
            const double dx = qx - px, dy = qy - py;
            if (ezm::ez_abs(dx) <= kEps || (dx * dx + dy * dy) <= kEps * kEps) {
                o.res_degen = true;
                o.jac_degen = true;
                o.emit[0] = false;
                return;
            }
            o.res[0] = ay - py - dy * (1.0 / dx) * (ax - px) - p0;
            if (JAC) {
                const double ipq = 1.0 / (px - qx);
                const double ip2 = ezm::ez_pow_m2(px - qx);
                o.pd[0][0] = (-py + qy) * ipq;
                o.pd[0][1] = 1.0;
                o.pd[0][2] = (ax - qx) * (py - qy) * ip2;
                o.pd[0][3] = (-ax + qx) * ipq;
                o.pd[0][4] = -(ax - px) * (py - qy) * ip2;
                o.pd[0][5] = (ax - px) * ipq;
            }
        } break;
        case EZPZ_K_HORIZONTAL_POINT_LINE_DISTANCE: {
            const double ax = x(d[0]), ay = x(d[1]), px = x(d[2]), py = x(d[3]), qx = x(d[4]), qy = x(d[5]);
            const double dx = qx - px, dy = qy - py;
            const double len2 = dx * dx + dy * dy;
            // residual tests '<=' (constraints.rs:778), the Jacobian '<' (:1750)
            if (ezm::ez_abs(dy) <= kEps || len2 <= kEps * kEps) {
                o.res_degen = true;
            } else {
                o.res[0] = ax - px - dx * (1.0 / dy) * (ay - py) - p0;
            }
            if (JAC) {
                if (ezm::ez_abs(dy) < kEps || len2 < kEps * kEps) {
                    o.jac_degen = true;
                    o.emit[0] = false;
                } else {
                    const double ipq = 1.0 / (py - qy);
                    const double ip2 = ezm::ez_pow_m2(py - qy);
                    o.pd[0][0] = 1.0;
                    o.pd[0][1] = (-px + qx) * ipq;
                    o.pd[0][2] = (-ay + qy) * ipq;
                    o.pd[0][3] = (ay - qy) * (px - qx) * ip2;
                    o.pd[0][4] = (ay - py) * ipq;
                    o.pd[0][5] = -(ay - py) * (px - qx) * ip2;
                }
            }
        } break;
        case EZPZ_K_SYMMETRIC: {
            const double px = x(d[0]), py = x(d[1]), qx = x(d[2]), qy = x(d[3]);
            const double ax = x(d[4]), ay = x(d[5]), bx = x(d[6]), by = x(d[7]);
            {  // reflect(a - p, q - p) - b + p   (vector.rs:58-69; unguarded division)
                const V2 s{ax - px, ay - py};
                const V2 l{qx - px, qy - py};
                const double f = vdot(s, l) / vdot(l, l);
                const V2 proj = vscale(l, f);
                const V2 rej = vsub(s, proj);
                const V2 refl = vsub(s, vscale(rej, 2.0));
                o.res[0] = (refl.x - bx) + px;
                o.res[1] = (refl.y - by) + py;
            }
            if (JAC) {  // pds_from_symmetric (constraints.rs:2361-2433)
                const double dx = px - qx, dy = py - qy;
                const double dx2 = dx * dx, dy2 = dy * dy;
                const double r = dx2 + dy2;
                const double r2 = ezm::ez_pow2(r);
                if (r2 < kEps) {
                    o.jac_degen = true;
                    o.emit[0] = false;
                    o.emit[1] = false;
                } else {
                    const double sx = ax - px, sy = ay - py;
                    const double dt = sx * dx + sy * dy;
                    o.pd[0][0] = (-4.0 * dx2 * dt + 2.0 * r2 + 2.0 * r * (sx * dx + sy * dy + dx * (ax - 2.0 * px + qx))) / r2;
                    o.pd[1][0] = dy * (-4.0 * dx * dt + 2.0 * r * (ax - 2.0 * px + qx)) / r2;
                    o.pd[0][1] = dx * (-4.0 * dy * dt + 2.0 * r * (ay - 2.0 * py + qy)) / r2;
                    o.pd[1][1] = (-4.0 * dy2 * dt + 2.0 * r2 + 2.0 * r * (sx * dx + sy * dy + dy * (ay - 2.0 * py + qy))) / r2;
                    o.pd[0][2] = (4.0 * dx2 * dt - (4.0 * sx * dx + 2.0 * sy * dy) * r) / r2;
                    o.pd[1][2] = dy * (-2.0 * sx * r + 4.0 * dx * dt) / r2;
                    o.pd[0][3] = dx * (-2.0 * sy * r + 4.0 * dy * dt) / r2;
                    o.pd[1][3] = (4.0 * dy2 * dt - (2.0 * sx * dx + 4.0 * sy * dy) * r) / r2;
                    o.pd[0][4] = 1.0 * (dx2 - dy2) / r;
                    o.pd[1][4] = 2.0 * dx * dy / r;
                    o.pd[0][5] = 2.0 * dx * dy / r;
                    o.pd[1][5] = 1.0 * (-dx2 + dy2) / r;
                    o.pd[0][6] = -1.0;
                    o.pd[1][6] = 0.0;
                    o.pd[0][7] = 0.0;
                    o.pd[1][7] = -1.0;
                }
            }
        } break;
        case EZPZ_K_POINT_ARC_COINCIDENT: {
            const V2 ce{x(d[4]), x(d[5])};
            const V2 s = vsub(V2{x(d[0]), x(d[1])}, ce);
            const V2 e = vsub(V2{x(d[2]), x(d[3])}, ce);
            const V2 p = vsub(V2{x(d[6]), x(d[7])}, ce);
            const double r = vmag(s), r_e = vmag(e), r_p = vmag(p);
            if (r < kEps || r_e < kEps || r_p < kEps) {
                o.res_degen = true;
                o.jac_degen = true;
                o.emit[0] = false;
                o.emit[1] = false;
                return;
            }
            const V2 e_proj = vscale(e, r / r_e);
            const int part = pac_classify(s, e_proj, p);
            V2 f;
            if (part == 0) f = vscale(p, r / r_p - 1.0);
            else if (part == 1) f = vsub(e_proj, p);
            else f = vsub(s, p);
            o.res[0] = f.x;
            o.res[1] = f.y;
            if (JAC) {
                const V2 u_s = vscale(s, 1.0 / r);
                const V2 u_e = vscale(e, 1.0 / r_e);
                double js00, js01, js10, js11, je00, je01, je10, je11, jp00, jp01, jp10, jp11;
                if (part == 0) {
                    const V2 u_p = vscale(p, 1.0 / r_p);
                    const double q = r / r_p;
                    js00 = u_p.x * u_s.x; js01 = u_p.y * u_s.x;
                    js10 = u_p.x * u_s.y; js11 = u_p.y * u_s.y;
                    je00 = je01 = je10 = je11 = 0.0;
                    jp00 = (q - 1.0) - q * u_p.x * u_p.x;
                    jp01 = -q * u_p.y * u_p.x;
                    jp10 = -q * u_p.x * u_p.y;
                    jp11 = (q - 1.0) - q * u_p.y * u_p.y;
                } else if (part == 1) {
                    const double q = r / r_e;
                    js00 = u_e.x * u_s.x; js01 = u_e.y * u_s.x;
                    js10 = u_e.x * u_s.y; js11 = u_e.y * u_s.y;
                    je00 = q * (1.0 - u_e.x * u_e.x);
                    je01 = -q * u_e.y * u_e.x;
                    je10 = -q * u_e.x * u_e.y;
                    je11 = q * (1.0 - u_e.y * u_e.y);
                    jp00 = -1.0; jp01 = 0.0; jp10 = 0.0; jp11 = -1.0;
                } else {
                    js00 = 1.0; js01 = 0.0; js10 = 0.0; js11 = 1.0;
                    je00 = je01 = je10 = je11 = 0.0;
                    jp00 = -1.0; jp01 = 0.0; jp10 = 0.0; jp11 = -1.0;
                }
                const double jo00 = -(js00 + je00 + jp00), jo01 = -(js01 + je01 + jp01);
                const double jo10 = -(js10 + je10 + jp10), jo11 = -(js11 + je11 + jp11);
                // emission order: cx, cy, sx, sy, ex, ey, px, py
                o.pd[0][0] = jo00; o.pd[0][1] = jo10; o.pd[0][2] = js00; o.pd[0][3] = js10;
                o.pd[0][4] = je00; o.pd[0][5] = je10; o.pd[0][6] = jp00; o.pd[0][7] = jp10;
                o.pd[1][0] = jo01; o.pd[1][1] = jo11; o.pd[1][2] = js01; o.pd[1][3] = js11;
                o.pd[1][4] = je01; o.pd[1][5] = je11; o.pd[1][6] = jp01; o.pd[1][7] = jp11;
            }
        } break;
        case EZPZ_K_ARC_LENGTH: {
            const double ax = x(d[0]), ay = x(d[1]), bx = x(d[2]), by = x(d[3]), cx = x(d[4]), cy = x(d[5]);
            const double ux = ax - cx, uy = ay - cy;
            const double r2 = ux * ux + uy * uy;
            if (r2 <= kEps * kEps) {
                o.res_degen = true;
                o.jac_degen = true;
                o.emit[0] = false;
                o.emit[1] = false;
                return;
            }
            const double r = EZ_SQRT(r2);
            const double alpha = p0 / r;
            double sa, ca;
            ezm::ez_sincos(alpha, sa, ca);
            const double rux = ca * ux - sa * uy;
            const double ruy = sa * ux + ca * uy;
            o.res[0] = (bx - cx) - rux;
            o.res[1] = (by - cy) - ruy;
            if (JAC) {
                const double k = p0 / (r2 * r);
                o.pd[0][0] = -ca - ruy * ux * k;
                o.pd[0][1] = sa - ruy * uy * k;
                o.pd[0][2] = 1.0;
                o.pd[0][3] = 0.0;
                o.pd[0][4] = -1.0 + ca + ruy * ux * k;
                o.pd[0][5] = -sa + ruy * uy * k;
                o.pd[1][0] = -sa + rux * ux * k;
                o.pd[1][1] = -ca + rux * uy * k;
                o.pd[1][2] = 0.0;
                o.pd[1][3] = 1.0;
                o.pd[1][4] = sa - rux * ux * k;
                o.pd[1][5] = -1.0 + ca - rux * uy * k;
            }
        } break;
        case EZPZ_K_ARC_ANGLE:
            lines_at_angle<JAC>(x(d[4]), x(d[5]), x(d[0]), x(d[1]), x(d[4]), x(d[5]), x(d[2]), x(d[3]), p0, p1, o);
            break;
        case EZPZ_K_POINTS_AT_ANGLE: {
            const V2 a{x(d[0]), x(d[1])}, b{x(d[2]), x(d[3])}, c2{x(d[4]), x(d[5])};
            const V2 u = vsub(b, a), v = vsub(c2, a);
            const double len_u = vmag(u), len_v = vmag(v);
            if (len_u <= kEps || len_v <= kEps) {
                o.res_degen = true;
                o.jac_degen = true;
                o.emit[0] = false;
                o.emit[1] = false;
                return;
            }
            const double sden = (len_u + len_v) * 0.5;
            const V2 rot_u = rot_apply(p0, p1, u);
            const double inv_s = 1.0 / sden;  // residual() writes `* (1.0 / s)`, jacobian_rows `* inv_s`: same value
            const V2 res = vscale(vsub(vscale(v, len_u), vscale(rot_u, len_v)), inv_s);
            o.res[0] = res.x;
            o.res[1] = res.y;
            if (JAC) {
                const V2 u_hat = vscale(u, 1.0 / len_u);
                const V2 v_hat = vscale(v, 1.0 / len_v);
                const V2 re1 = rot_apply(p0, p1, V2{1.0, 0.0});
                const V2 re2 = rot_apply(p0, p1, V2{0.0, 1.0});
                const V2 half = vscale(res, 0.5);
                const V2 vmh = vsub(v, half);
                const V2 rph = vadd(rot_u, half);
                const V2 du0 = vscale(vsub(vscale(vmh, u_hat.x), vscale(re1, len_v)), inv_s);
                const V2 du1 = vscale(vsub(vscale(vmh, u_hat.y), vscale(re2, len_v)), inv_s);
                const V2 dv0 = vscale(vsub(V2{len_u, 0.0}, vscale(rph, v_hat.x)), inv_s);
                const V2 dv1 = vscale(vsub(V2{0.0, len_u}, vscale(rph, v_hat.y)), inv_s);
                o.pd[0][0] = -(du0.x + dv0.x);
                o.pd[0][1] = -(du1.x + dv1.x);
                o.pd[0][2] = du0.x;
                o.pd[0][3] = du1.x;
                o.pd[0][4] = dv0.x;
                o.pd[0][5] = dv1.x;
                o.pd[1][0] = -(du0.y + dv0.y);
                o.pd[1][1] = -(du1.y + dv1.y);
                o.pd[1][2] = du0.y;
                o.pd[1][3] = du1.y;
                o.pd[1][4] = dv0.y;
                o.pd[1][5] = dv1.y;
            }
        } break;
        default: break;
    }
}

// Constraint::set_from_initial_values (constraints.rs:146-193): resolve an Undefined side from x.
template <class X>
EZ_HD uint32_t resolve_side(uint32_t kind, uint32_t side, const uint32_t* __restrict__ d, const X& x) {
    if (side != EZPZ_SIDE_UNDEFINED) return side;
    if (kind == EZPZ_K_LINE_TANGENT_TO_CIRCLE) {
        const V2 a{x(d[0]), x(d[1])}, b{x(d[2]), x(d[3])}, ce{x(d[4]), x(d[5])};
        return (vcross(vsub(b, a), vsub(ce, a)) >= 0.0) ? EZPZ_LINE_SIDE_LEFT : EZPZ_LINE_SIDE_RIGHT;
    }
    if (kind == EZPZ_K_CIRCLE_TANGENT_TO_CIRCLE) {
        const V2 ac{x(d[0]), x(d[1])}, bc{x(d[3]), x(d[4])};
        const double ar = x(d[2]), br = x(d[5]);
        const double dist = vmag(vsub(ac, bc));
        const double r_int = ezm::ez_abs(ezm::ez_abs(ar - br) - dist);
        const double r_ext = ezm::ez_abs(ar + br - dist);
        return (r_int < r_ext) ? EZPZ_CIRCLE_SIDE_INTERIOR : EZPZ_CIRCLE_SIDE_EXTERIOR;
    }
    return side;
}

}  // namespace ezd
