// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under ezpz_b200/ may include, link or call this.
//
// Per-constraint sparsity lists, residuals and analytic Jacobian rows: a CPU restatement of
// ezpz/src/constraints.rs (`nonzeroes` :378-491, `residual` :499-950, `residual_dim` :954-993,
// `jacobian_rows` :1000-2293, helpers :2361-2647) and ezpz/src/vector.rs.  Written as plain
// sequential C++ in the reference's own evaluation order: Rust never contracts a*b+c, so this file
// must be compiled with -ffp-contract=off.
//
// The record layout mirrors include/ezpz_b200.h's ezpz_constraint_t (the oracle declares its own
// copy so that it does not depend on the product tree).
#pragma once
#include <cstdint>
#include <vector>

#include "libm_port.h"

namespace orc {

struct Rec {
    uint32_t kind;
    uint32_t flags;
    uint32_t ids[8];
    double p0;
    double p1;
    double weight;
};
static_assert(sizeof(Rec) == 64, "record must be 64 bytes");

enum Kind : uint32_t {
    K_LTC = 0, K_CTC, K_DISTANCE, K_DISTANCE_VAR, K_VDIST, K_HDIST, K_VERTICAL, K_HORIZONTAL,
    K_LINES_AT_ANGLE, K_FIXED, K_SCALAR_EQUAL, K_POINTS_COINCIDENT, K_CIRCLE_RADIUS,
    K_LINES_EQUAL_LENGTH, K_ARC_RADIUS, K_ARC, K_MIDPOINT, K_PLD, K_VPLD, K_HPLD, K_SYMMETRIC,
    K_PAC, K_ARC_LENGTH, K_ARC_ANGLE, K_POINTS_AT_ANGLE, K_COUNT
};
enum : uint32_t { SIDE_UNDEFINED = 0, LINE_LEFT = 1, LINE_RIGHT = 2, CIRCLE_EXTERIOR = 1, CIRCLE_INTERIOR = 2 };

constexpr double EPSILON = 1e-4;  // lib.rs:43

// ---- vector.rs
struct V {
    double x, y;
};
static inline V operator-(V a, V b) { return {a.x - b.x, a.y - b.y}; }
static inline V operator+(V a, V b) { return {a.x + b.x, a.y + b.y}; }
static inline V operator*(V a, double s) { return {a.x * s, a.y * s}; }
static inline double magnitude(V a) { return orc_hypot(a.x, a.y); }                   // :15-17
static inline double magnitude_squared(V a) { return orc_pow2(a.x) + orc_pow2(a.y); } // :20-22
static inline double dot(V a, V b) { return a.x * b.x + a.y * b.y; }                  // :25-27
static inline double euclidean_distance(V a, V b) { return magnitude(a - b); }        // :30-33
static inline double cross_2d(V a, V b) { return a.x * b.y - a.y * b.x; }             // :37-39
static inline V perp_ccw(V a) { return {-a.y, a.x}; }
static inline V perp_cw(V a) { return {a.y, -a.x}; }
static inline V project(V a, V b) { return b * (dot(a, b) / dot(b, b)); }             // :58-60
static inline V reject(V a, V b) { return a - project(a, b); }                        // :63-65
static inline V reflect(V a, V b) { return a - (reject(a, b) * 2.0); }                // :67-69
static inline double signed_angle(V a, V b) { return orc_atan2(cross_2d(a, b), dot(a, b)); }  // :72-74

struct Rot {  // Rotation2, :105-143; col0 = (cos, sin)
    double c, s;
};
static inline V apply(Rot r, V v) { return {(r.c * v.x) - (r.s * v.y), (r.s * v.x) + (r.c * v.y)}; }
static inline Rot inverse(Rot r) { return {r.c, -r.s}; }

// constraints.rs:954-993
static inline int residual_dim(const Rec& c) {
    switch (c.kind) {
        case K_POINTS_COINCIDENT:
        case K_ARC_RADIUS:
        case K_MIDPOINT:
        case K_SYMMETRIC:
        case K_PAC:
        case K_ARC_LENGTH:
        case K_POINTS_AT_ANGLE:
            return 2;
        default:
            return 1;
    }
}

// constraints.rs:378-491.  Lists may contain duplicates; order is the reference's.
static inline void nonzeroes(const Rec& c, std::vector<uint32_t>& row0, std::vector<uint32_t>& row1) {
    const uint32_t* d = c.ids;
    auto ext = [](std::vector<uint32_t>& r, std::initializer_list<uint32_t> l) { r.insert(r.end(), l); };
    switch (c.kind) {
        case K_LTC: ext(row0, {d[0], d[1], d[2], d[3], d[4], d[5], d[6]}); break;
        case K_CTC: ext(row0, {d[0], d[1], d[2], d[3], d[4], d[5]}); break;
        case K_DISTANCE: ext(row0, {d[0], d[1], d[2], d[3]}); break;
        case K_DISTANCE_VAR: ext(row0, {d[0], d[1], d[2], d[3], d[4]}); break;
        case K_VDIST: ext(row0, {d[1], d[3]}); break;
        case K_HDIST: ext(row0, {d[0], d[2]}); break;
        case K_VERTICAL: ext(row0, {d[0], d[2]}); break;
        case K_HORIZONTAL: ext(row0, {d[1], d[3]}); break;
        case K_LINES_AT_ANGLE: ext(row0, {d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]}); break;
        case K_FIXED: ext(row0, {d[0]}); break;
        case K_SCALAR_EQUAL: ext(row0, {d[0], d[1]}); break;
        case K_POINTS_COINCIDENT:
            ext(row0, {d[0], d[2]});
            ext(row1, {d[1], d[3]});
            break;
        case K_CIRCLE_RADIUS: ext(row0, {d[2]}); break;
        case K_LINES_EQUAL_LENGTH: ext(row0, {d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]}); break;
        case K_ARC_RADIUS:  // Distance(center,start) into row0, Distance(center,end) into row1 (:422-431)
            ext(row0, {d[4], d[5], d[0], d[1]});
            ext(row1, {d[4], d[5], d[2], d[3]});
            break;
        case K_ARC: ext(row0, {d[0], d[1], d[2], d[3], d[4], d[5]}); break;
        case K_MIDPOINT:
            ext(row0, {d[0], d[2], d[4]});
            ext(row1, {d[1], d[3], d[5]});
            break;
        case K_PLD: ext(row0, {d[0], d[1], d[2], d[3], d[4], d[5]}); break;
        case K_VPLD:
        case K_HPLD: ext(row0, {d[2], d[3], d[4], d[5], d[0], d[1]}); break;
        case K_SYMMETRIC:
        case K_PAC:
            ext(row0, {d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]});
            ext(row1, {d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]});
            break;
        case K_ARC_LENGTH:
            ext(row0, {d[0], d[1], d[2], d[3], d[4], d[5]});
            ext(row1, {d[0], d[1], d[2], d[3], d[4], d[5]});
            break;
        case K_ARC_ANGLE:  // LinesAtAngle(center->start, center->end) (:470-481)
            ext(row0, {d[4], d[5], d[0], d[1], d[4], d[5], d[2], d[3]});
            break;
        case K_POINTS_AT_ANGLE:
            ext(row0, {d[0], d[1], d[2], d[3], d[4], d[5]});
            ext(row1, {d[0], d[1], d[2], d[3], d[4], d[5]});
            break;
        default: break;
    }
}

// constraints.rs:146-193 — resolve Undefined sides from values indexed BY ID.
// constraints.rs:146-193.  `n_values` = length of the id-indexed table (max guessed id + 1, lib.rs:172-178).  The reference
// indexes it unchecked and would PANIC on an id beyond it (which its own fuzz target forbids); such ids read 0.0 here and the
// solve then fails with MissingGuess in validate_variables, as it does for every other kind.
static inline void set_from_initial_values(Rec& c, const double* values, size_t n_values = (size_t)-1) {
    auto iv = [&](uint32_t id) { return (size_t)id < n_values ? values[id] : 0.0; };
    if (c.kind == K_LTC && c.flags == SIDE_UNDEFINED) {
        V p0{iv(c.ids[0]), iv(c.ids[1])}, p1{iv(c.ids[2]), iv(c.ids[3])}, ce{iv(c.ids[4]), iv(c.ids[5])};
        c.flags = (cross_2d(p1 - p0, ce - p0) >= 0.0) ? LINE_LEFT : LINE_RIGHT;
    } else if (c.kind == K_CTC && c.flags == SIDE_UNDEFINED) {
        V a_c{iv(c.ids[0]), iv(c.ids[1])};
        double a_r = iv(c.ids[2]);
        V b_c{iv(c.ids[3]), iv(c.ids[4])};
        double b_r = iv(c.ids[5]);
        double dist = magnitude(a_c - b_c);
        double r_int = std::fabs(std::fabs(a_r - b_r) - dist);
        double r_ext = std::fabs(a_r + b_r - dist);
        c.flags = (r_int < r_ext) ? CIRCLE_INTERIOR : CIRCLE_EXTERIOR;
    }
}

enum PacPart { PAC_INTERIOR, PAC_START, PAC_END };
// constraints.rs:2593-2606
static inline PacPart classify_point_arc_coincident(V s, V e, V p) {
    const double two_pi = 2.0 * 3.14159265358979323846;
    double a_sp = orc_rem_euclid(signed_angle(s, p), two_pi);
    double a_se = orc_rem_euclid(signed_angle(s, e), two_pi);
    if (a_sp < a_se) return PAC_INTERIOR;
    if (magnitude_squared(e - p) < magnitude_squared(s - p)) return PAC_END;
    return PAC_START;
}

static void lines_at_angle_residual(double x0, double y0, double x1, double y1, double x2, double y2,
                                    double x3, double y3, Rot rot, double* r0, bool* degenerate) {
    V u{x1 - x0, y1 - y0};
    V v{x3 - x2, y3 - y2};
    double len_u = magnitude(u), len_v = magnitude(v);
    if (len_u <= EPSILON || len_v <= EPSILON) {
        *degenerate = true;
        return;
    }
    *r0 = cross_2d(u, apply(inverse(rot), v)) / ((len_u + len_v) * 0.5);
}

// constraints.rs:499-950.  r0/r1 are pre-set to 0 by the caller (solver.rs:326-329).
static inline void residual(const Rec& c, const double* x, double* r0, double* r1, bool* degenerate) {
    const uint32_t* d = c.ids;
    switch (c.kind) {
        case K_LTC: {
            V p0{x[d[0]], x[d[1]]}, p1{x[d[2]], x[d[3]]}, ce{x[d[4]], x[d[5]]};
            double radius = std::fabs(x[d[6]]);
            V u = p1 - p0;
            double mag_u = magnitude(u);
            if (mag_u <= EPSILON) {
                *r0 = 0.0;
                *degenerate = true;
                return;
            }
            V v = ce - p0;
            double cross_uv = cross_2d(u, v);
            double side_sign = (c.flags == LINE_RIGHT) ? -1.0 : 1.0;
            double cen_dist = side_sign * cross_uv / mag_u;
            *r0 = cen_dist - radius;
        } break;
        case K_CTC: {
            V a_c{x[d[0]], x[d[1]]};
            double a_r = std::fabs(x[d[2]]);
            V b_c{x[d[3]], x[d[4]]};
            double b_r = std::fabs(x[d[5]]);
            double dist = magnitude(a_c - b_c);
            *r0 = (c.flags == CIRCLE_INTERIOR) ? std::fabs(a_r - b_r) - dist : a_r + b_r - dist;
        } break;
        case K_DISTANCE: {
            V p0{x[d[0]], x[d[1]]}, p1{x[d[2]], x[d[3]]};
            *r0 = euclidean_distance(p0, p1) - c.p0;
        } break;
        case K_DISTANCE_VAR: {
            double px = x[d[0]], py = x[d[1]], qx = x[d[2]], qy = x[d[3]], dd = x[d[4]];
            *r0 = -dd + std::sqrt(orc_pow2(px - qx) + orc_pow2(py - qy));
        } break;
        case K_VDIST: *r0 = (x[d[1]] - x[d[3]]) - c.p0; break;
        case K_HDIST: *r0 = (x[d[0]] - x[d[2]]) - c.p0; break;
        case K_VERTICAL: *r0 = x[d[0]] - x[d[2]]; break;
        case K_HORIZONTAL: *r0 = x[d[1]] - x[d[3]]; break;
        case K_FIXED: *r0 = x[d[0]] - c.p0; break;
        case K_SCALAR_EQUAL: *r0 = x[d[0]] - x[d[1]]; break;
        case K_LINES_AT_ANGLE:
            lines_at_angle_residual(x[d[0]], x[d[1]], x[d[2]], x[d[3]], x[d[4]], x[d[5]], x[d[6]], x[d[7]],
                                    Rot{c.p0, c.p1}, r0, degenerate);
            break;
        case K_POINTS_COINCIDENT:
            *r0 = x[d[0]] - x[d[2]];
            *r1 = x[d[1]] - x[d[3]];
            break;
        case K_CIRCLE_RADIUS: *r0 = x[d[2]] - c.p0; break;
        case K_LINES_EQUAL_LENGTH: {
            V a0{x[d[0]], x[d[1]]}, a1{x[d[2]], x[d[3]]}, b0{x[d[4]], x[d[5]]}, b1{x[d[6]], x[d[7]]};
            double len0 = euclidean_distance(a0, a1);
            double len1 = euclidean_distance(b0, b1);
            *r0 = len0 - len1;
        } break;
        case K_ARC_RADIUS: {  // two Distance(center, .) residuals (:659-682)
            V s{x[d[0]], x[d[1]]}, e{x[d[2]], x[d[3]]}, ce{x[d[4]], x[d[5]]};
            *r0 = euclidean_distance(ce, s) - c.p0;
            *r1 = euclidean_distance(ce, e) - c.p0;
        } break;
        case K_ARC: {
            double sx = x[d[0]], sy = x[d[1]], ex = x[d[2]], ey = x[d[3]], cx = x[d[4]], cy = x[d[5]];
            double dist0 = orc_hypot(sx - cx, sy - cy);
            double dist1 = orc_hypot(ex - cx, ey - cy);
            *r0 = dist0 - dist1;
        } break;
        case K_MIDPOINT: {
            double px = x[d[0]], py = x[d[1]], qx = x[d[2]], qy = x[d[3]], ax = x[d[4]], ay = x[d[5]];
            *r0 = ax - px / 2.0 - qx / 2.0;
            *r1 = ay - py / 2.0 - qy / 2.0;
        } break;
        case K_PLD: {
            double ptx = x[d[0]], pty = x[d[1]];
            double px = x[d[2]], py = x[d[3]], qx = x[d[4]], qy = x[d[5]];
            double a = py - qy, b = qx - px, cc = (px * qy) - (qx * py);  // :2625-2639
            double denominator = orc_hypot(a, b);
            if (denominator < EPSILON) {
                *r0 = 0.0;
                *degenerate = true;
                return;
            }
            double actual = (a * ptx + b * pty + cc) / denominator;
            *r0 = actual - c.p0;
        } break;
        case K_VPLD: {
            double ax = x[d[0]], ay = x[d[1]], px = x[d[2]], py = x[d[3]], qx = x[d[4]], qy = x[d[5]];
            double dx = qx - px, dy = qy - py;
            if (std::fabs(dx) <= EPSILON || (dx * dx + dy * dy) <= EPSILON * EPSILON) {
                *degenerate = true;
                return;
            }
            *r0 = ay - py - dy * (1.0 / dx) * (ax - px) - c.p0;
        } break;
        case K_HPLD: {
            double ax = x[d[0]], ay = x[d[1]], px = x[d[2]], py = x[d[3]], qx = x[d[4]], qy = x[d[5]];
            double dx = qx - px, dy = qy - py;
            if (std::fabs(dy) <= EPSILON || (dx * dx + dy * dy) <= EPSILON * EPSILON) {
                *degenerate = true;
                return;
            }
            *r0 = ax - px - dx * (1.0 / dy) * (ay - py) - c.p0;
        } break;
        case K_SYMMETRIC: {
            V p{x[d[0]], x[d[1]]}, q{x[d[2]], x[d[3]]}, a{x[d[4]], x[d[5]]}, b{x[d[6]], x[d[7]]};
            V res = reflect(a - p, q - p) - b + p;
            *r0 = res.x;
            *r1 = res.y;
        } break;
        case K_PAC: {
            V ce{x[d[4]], x[d[5]]};
            V s = V{x[d[0]], x[d[1]]} - ce;
            V e = V{x[d[2]], x[d[3]]} - ce;
            V p = V{x[d[6]], x[d[7]]} - ce;
            double r = magnitude(s), r_e = magnitude(e), r_p = magnitude(p);
            if (r < EPSILON || r_e < EPSILON || r_p < EPSILON) {
                *r0 = 0.0;
                *r1 = 0.0;
                *degenerate = true;
                return;
            }
            V e_proj = e * (r / r_e);
            V f;
            switch (classify_point_arc_coincident(s, e_proj, p)) {
                case PAC_INTERIOR: f = p * (r / r_p - 1.0); break;
                case PAC_END: f = e_proj - p; break;
                default: f = s - p; break;
            }
            *r0 = f.x;
            *r1 = f.y;
        } break;
        case K_ARC_LENGTH: {
            double ax = x[d[0]], ay = x[d[1]], bx = x[d[2]], by = x[d[3]], cx = x[d[4]], cy = x[d[5]];
            double ux = ax - cx, uy = ay - cy;
            double r2 = ux * ux + uy * uy;
            if (r2 <= EPSILON * EPSILON) {
                *r0 = 0.0;
                *r1 = 0.0;
                *degenerate = true;
                return;
            }
            double alpha = c.p0 / std::sqrt(r2);
            double sa = orc_sin(alpha), ca = orc_cos(alpha);
            double rux = ca * ux - sa * uy;
            double ruy = sa * ux + ca * uy;
            *r0 = (bx - cx) - rux;
            *r1 = (by - cy) - ruy;
        } break;
        case K_ARC_ANGLE:  // LinesAtAngle(center->start, center->end, Other) (:897-915)
            lines_at_angle_residual(x[d[4]], x[d[5]], x[d[0]], x[d[1]], x[d[4]], x[d[5]], x[d[2]], x[d[3]],
                                    Rot{c.p0, c.p1}, r0, degenerate);
            break;
        case K_POINTS_AT_ANGLE: {
            V p0{x[d[0]], x[d[1]]}, p1{x[d[2]], x[d[3]]}, p2{x[d[4]], x[d[5]]};
            V u = p1 - p0, v = p2 - p0;
            double len_u = magnitude(u), len_v = magnitude(v);
            if (len_u <= EPSILON || len_v <= EPSILON) {
                *degenerate = true;
                return;
            }
            Rot rot{c.p0, c.p1};
            double s = (len_u + len_v) * 0.5;
            V res = (v * len_u - apply(rot, u) * len_v) * (1.0 / s);
            *r0 = res.x;
            *r1 = res.y;
        } break;
        default: break;
    }
}

struct JVar {
    uint32_t id;
    double pd;
};
typedef std::vector<JVar> JRow;

static void lines_at_angle_jacobian(const uint32_t idv[8], double x0, double y0, double x1, double y1,
                                    double x2, double y2, double x3, double y3, Rot rot, JRow& row0,
                                    bool* degenerate) {
    V u{x1 - x0, y1 - y0};
    V v{x3 - x2, y3 - y2};
    double len_u = magnitude(u), len_v = magnitude(v);
    if ((len_u <= EPSILON) || (len_v <= EPSILON)) {
        *degenerate = true;
        return;
    }
    V u_hat = u * (1.0 / len_u);
    V v_hat = v * (1.0 / len_v);
    double s = (len_u + len_v) * 0.5;
    double a = cross_2d(u, apply(inverse(rot), v));
    double inv_s = 1.0 / s;
    double t = a * inv_s * 0.5;
    V df_du = (perp_cw(apply(inverse(rot), v)) - u_hat * t) * inv_s;
    V df_dv = (perp_ccw(apply(rot, u)) - v_hat * t) * inv_s;
    const double pds[8] = {-df_du.x, -df_du.y, df_du.x, df_du.y, -df_dv.x, -df_dv.y, df_dv.x, df_dv.y};
    for (int k = 0; k < 8; ++k) row0.push_back({idv[k], pds[k]});
}

static void distance_jacobian(uint32_t i0x, uint32_t i0y, uint32_t i1x, uint32_t i1y, const double* x,
                              JRow& row, bool* degenerate) {  // :1160-1204
    double x0 = x[i0x], y0 = x[i0y], x1 = x[i1x], y1 = x[i1y];
    double dist = euclidean_distance(V{x0, y0}, V{x1, y1});
    if (dist < EPSILON) {
        *degenerate = true;
        return;
    }
    row.push_back({i0x, (x0 - x1) / dist});
    row.push_back({i0y, (y0 - y1) / dist});
    row.push_back({i1x, (-x0 + x1) / dist});
    row.push_back({i1y, (-y0 + y1) / dist});
}

// constraints.rs:1000-2293.  Degenerate rows emit nothing.
static inline void jacobian_rows(const Rec& c, const double* x, JRow& row0, JRow& row1, bool* degenerate) {
    const uint32_t* d = c.ids;
    switch (c.kind) {
        case K_LTC: {
            V p0{x[d[0]], x[d[1]]}, p1{x[d[2]], x[d[3]]}, ce{x[d[4]], x[d[5]]};
            V u = p1 - p0;
            double mag_u = magnitude(u);
            if (mag_u <= EPSILON) {
                *degenerate = true;
                return;
            }
            V v = ce - p0;
            double cross_uv = cross_2d(u, v);
            double mag_u_cubed = mag_u * mag_u * mag_u;
            double side_sign = (c.flags == LINE_RIGHT) ? -1.0 : 1.0;
            double dr_du_x = side_sign * (-(u.x * cross_uv) / mag_u_cubed + v.y / mag_u);
            double dr_du_y = side_sign * (-(u.y * cross_uv) / mag_u_cubed - v.x / mag_u);
            double dr_dv_x = side_sign * (-u.y / mag_u);
            double dr_dv_y = side_sign * (u.x / mag_u);
            double radius = x[d[6]];
            double dr_dr = -orc_signum(radius);
            row0.push_back({d[0], -(dr_du_x + dr_dv_x)});
            row0.push_back({d[1], -(dr_du_y + dr_dv_y)});
            row0.push_back({d[2], dr_du_x});
            row0.push_back({d[3], dr_du_y});
            row0.push_back({d[4], dr_dv_x});
            row0.push_back({d[5], dr_dv_y});
            row0.push_back({d[6], dr_dr});
        } break;
        case K_CTC: {
            V a_c{x[d[0]], x[d[1]]};
            double a_r = x[d[2]];
            V b_c{x[d[3]], x[d[4]]};
            double b_r = x[d[5]];
            V dd = b_c - a_c;
            double mag_d = magnitude(dd);
            if (mag_d <= EPSILON) {
                *degenerate = true;
                return;
            }
            V u_d = dd * (1.0 / mag_d);
            double a_sign = orc_signum(a_r), b_sign = orc_signum(b_r);
            double dr_dar, dr_dbr;
            if (c.flags == CIRCLE_INTERIOR) {
                double inner = orc_signum(std::fabs(a_r) - std::fabs(b_r));
                dr_dar = inner * a_sign;
                dr_dbr = -inner * b_sign;
            } else {
                dr_dar = a_sign;
                dr_dbr = b_sign;
            }
            row0.push_back({d[0], u_d.x});
            row0.push_back({d[1], u_d.y});
            row0.push_back({d[2], dr_dar});
            row0.push_back({d[3], -u_d.x});
            row0.push_back({d[4], -u_d.y});
            row0.push_back({d[5], dr_dbr});
        } break;
        case K_DISTANCE: distance_jacobian(d[0], d[1], d[2], d[3], x, row0, degenerate); break;
        case K_DISTANCE_VAR: {
            double px = x[d[0]], py = x[d[1]], qx = x[d[2]], qy = x[d[3]];
            double dist = euclidean_distance(V{px, py}, V{qx, qy});
            if (dist < EPSILON) {
                *degenerate = true;
                return;
            }
            row0.push_back({d[0], (px - qx) * (1.0 / dist)});
            row0.push_back({d[1], (py - qy) * (1.0 / dist)});
            row0.push_back({d[2], -(px - qx) * (1.0 / dist)});
            row0.push_back({d[3], -(py - qy) * (1.0 / dist)});
            row0.push_back({d[4], -1.0});
        } break;
        case K_VDIST:
            row0.push_back({d[1], 1.0});
            row0.push_back({d[3], -1.0});
            break;
        case K_HDIST:
            row0.push_back({d[0], 1.0});
            row0.push_back({d[2], -1.0});
            break;
        case K_VERTICAL:
            row0.push_back({d[0], 1.0});
            row0.push_back({d[2], -1.0});
            break;
        case K_HORIZONTAL:
            row0.push_back({d[1], 1.0});
            row0.push_back({d[3], -1.0});
            break;
        case K_FIXED: row0.push_back({d[0], 1.0}); break;
        case K_SCALAR_EQUAL:
            row0.push_back({d[0], 1.0});
            row0.push_back({d[1], -1.0});
            break;
        case K_LINES_AT_ANGLE:
            lines_at_angle_jacobian(d, x[d[0]], x[d[1]], x[d[2]], x[d[3]], x[d[4]], x[d[5]], x[d[6]], x[d[7]],
                                    Rot{c.p0, c.p1}, row0, degenerate);
            break;
        case K_POINTS_COINCIDENT:
            row0.push_back({d[0], 1.0});
            row0.push_back({d[2], -1.0});
            row1.push_back({d[1], 1.0});
            row1.push_back({d[3], -1.0});
            break;
        case K_CIRCLE_RADIUS: row0.push_back({d[2], 1.0}); break;
        case K_LINES_EQUAL_LENGTH: {
            double x0 = x[d[0]], y0 = x[d[1]], x1 = x[d[2]], y1 = x[d[3]];
            double x2 = x[d[4]], y2 = x[d[5]], x3 = x[d[6]], y3 = x[d[7]];
            double len0 = euclidean_distance(V{x0, y0}, V{x1, y1});
            double len1 = euclidean_distance(V{x2, y2}, V{x3, y3});
            if (len0 < EPSILON || len1 < EPSILON) {
                *degenerate = true;
                return;
            }
            const double pds[8] = {(x0 - x1) / len0,  (y0 - y1) / len0,  (-x0 + x1) / len0, (-y0 + y1) / len0,
                                   (-x2 + x3) / len1, (-y2 + y3) / len1, (x2 - x3) / len1,  (y2 - y3) / len1};
            for (int k = 0; k < 8; ++k) row0.push_back({d[k], pds[k]});
        } break;
        case K_ARC_RADIUS:  // each sub-call may independently flag degenerate and emit nothing (:1513-1536)
            distance_jacobian(d[4], d[5], d[0], d[1], x, row0, degenerate);
            distance_jacobian(d[4], d[5], d[2], d[3], x, row1, degenerate);
            break;
        case K_ARC: {
            double sx = x[d[0]], sy = x[d[1]], ex = x[d[2]], ey = x[d[3]], cx = x[d[4]], cy = x[d[5]];
            double usx = sx - cx, usy = sy - cy, uex = ex - cx, uey = ey - cy;
            double dist0 = orc_hypot(usx, usy), dist1 = orc_hypot(uex, uey);
            if (dist0 <= EPSILON || dist1 <= EPSILON) {
                *degenerate = true;
                return;
            }
            row0.push_back({d[0], usx / dist0});
            row0.push_back({d[1], usy / dist0});
            row0.push_back({d[2], -uex / dist1});
            row0.push_back({d[3], -uey / dist1});
            row0.push_back({d[4], -usx / dist0 + uex / dist1});
            row0.push_back({d[5], -usy / dist0 + uey / dist1});
        } break;
        case K_MIDPOINT:
            row0.push_back({d[4], 1.0});
            row0.push_back({d[0], -0.5});
            row0.push_back({d[2], -0.5});
            row1.push_back({d[5], 1.0});
            row1.push_back({d[1], -0.5});
            row1.push_back({d[3], -0.5});
            break;
        case K_PLD: {  // pds_for_point_line (:2435-2516); note: no degenerate guard
            double px = x[d[0]], py = x[d[1]], p0x = x[d[2]], p0y = x[d[3]], p1x = x[d[4]], p1y = x[d[5]];
            double euclid_dist = orc_hypot(-p0x + p1x, p0y - p1y);
            double d_px = (p0y - p1y) / euclid_dist;
            double d_py = (-p0x + p1x) / euclid_dist;
            double denom = orc_pow_1p5(orc_pow2(-p0x + p1x) + orc_pow2(p0y - p1y));
            double d_p0x = ((-p0x + p1x) * (p0x * p1y - p0y * p1x + px * (p0y - p1y) + py * (-p0x + p1x))) / denom +
                           (p1y - py) / euclid_dist;
            double d_p0y = ((-p0y + p1y) * (p0x * p1y - p0y * p1x + px * (p0y - p1y) + py * (-p0x + p1x))) / denom +
                           (-p1x + px) / euclid_dist;
            double d_p1x = ((p0x - p1x) * (p0x * p1y - p0y * p1x + px * (p0y - p1y) + py * (-p0x + p1x))) / denom +
                           (-p0y + py) / euclid_dist;
            double d_p1y = ((p0y - p1y) * (p0x * p1y - p0y * p1x + px * (p0y - p1y) + py * (-p0x + p1x))) / denom +
                           (p0x - px) / euclid_dist;
            row0.push_back({d[0], d_px});
            row0.push_back({d[1], d_py});
            row0.push_back({d[2], d_p0x});
            row0.push_back({d[3], d_p0y});
            row0.push_back({d[4], d_p1x});
            row0.push_back({d[5], d_p1y});
        } break;
        case K_VPLD: {
            double ax = x[d[0]], px = x[d[2]], py = x[d[3]], qx = x[d[4]], qy = x[d[5]];
            double dx = qx - px, dy = qy - py;
            if (std::fabs(dx) <= EPSILON || (dx * dx + dy * dy) <= EPSILON * EPSILON) {
                *degenerate = true;
                return;
            }
            double dpx = (ax - qx) * (py - qy) * orc_pow_m2(px - qx);
            double dpy = (-ax + qx) * (1.0 / (px - qx));
            double dqx = -(ax - px) * (py - qy) * orc_pow_m2(px - qx);
            double dqy = (ax - px) * (1.0 / (px - qx));
            double dax = (-py + qy) * (1.0 / (px - qx));
            row0.push_back({d[0], dax});
            row0.push_back({d[1], 1.0});
            row0.push_back({d[2], dpx});
            row0.push_back({d[3], dpy});
            row0.push_back({d[4], dqx});
            row0.push_back({d[5], dqy});
        } break;
        case K_HPLD: {
            double ay = x[d[1]], px = x[d[2]], py = x[d[3]], qx = x[d[4]], qy = x[d[5]];
            double dx = qx - px, dy = qy - py;
            if (std::fabs(dy) < EPSILON || (dx * dx + dy * dy) < EPSILON * EPSILON) {  // '<' here (:1750)
                *degenerate = true;
                return;
            }
            double dpx = (-ay + qy) * (1.0 / (py - qy));
            double dpy = (ay - qy) * (px - qx) * orc_pow_m2(py - qy);
            double dqx = (ay - py) * (1.0 / (py - qy));
            double dqy = -(ay - py) * (px - qx) * orc_pow_m2(py - qy);
            double day = (-px + qx) * (1.0 / (py - qy));
            row0.push_back({d[0], 1.0});
            row0.push_back({d[1], day});
            row0.push_back({d[2], dpx});
            row0.push_back({d[3], dpy});
            row0.push_back({d[4], dqx});
            row0.push_back({d[5], dqy});
        } break;
        case K_SYMMETRIC: {  // pds_from_symmetric (:2361-2433)
            double px = x[d[0]], py = x[d[1]], qx = x[d[2]], qy = x[d[3]], ax = x[d[4]], ay = x[d[5]];
            double dx = px - qx, dy = py - qy;
            double dx2 = dx * dx, dy2 = dy * dy;
            double r = dx2 + dy2;
            double r2 = orc_pow2(r);
            if (r2 < EPSILON) {
                *degenerate = true;
                return;
            }
            double sx = ax - px, sy = ay - py;
            double dt = sx * dx + sy * dy;
            double dpx0 = (-4.0 * dx2 * dt + 2.0 * r2 + 2.0 * r * (sx * dx + sy * dy + dx * (ax - 2.0 * px + qx))) / r2;
            double dpx1 = dy * (-4.0 * dx * dt + 2.0 * r * (ax - 2.0 * px + qx)) / r2;
            double dpy0 = dx * (-4.0 * dy * dt + 2.0 * r * (ay - 2.0 * py + qy)) / r2;
            double dpy1 = (-4.0 * dy2 * dt + 2.0 * r2 + 2.0 * r * (sx * dx + sy * dy + dy * (ay - 2.0 * py + qy))) / r2;
            double dqx0 = (4.0 * dx2 * dt - (4.0 * sx * dx + 2.0 * sy * dy) * r) / r2;
            double dqx1 = dy * (-2.0 * sx * r + 4.0 * dx * dt) / r2;
            double dqy0 = dx * (-2.0 * sy * r + 4.0 * dy * dt) / r2;
            double dqy1 = (4.0 * dy2 * dt - (2.0 * sx * dx + 4.0 * sy * dy) * r) / r2;
            double dax0 = 1.0 * (dx2 - dy2) / r, dax1 = 2.0 * dx * dy / r;
            double day0 = 2.0 * dx * dy / r, day1 = 1.0 * (-dx2 + dy2) / r;
            const double a0[8] = {dpx0, dpy0, dqx0, dqy0, dax0, day0, -1.0, 0.0};
            const double a1[8] = {dpx1, dpy1, dqx1, dqy1, dax1, day1, 0.0, -1.0};
            for (int k = 0; k < 8; ++k) {
                row0.push_back({d[k], a0[k]});
                row1.push_back({d[k], a1[k]});
            }
        } break;
        case K_PAC: {
            V ce{x[d[4]], x[d[5]]};
            V s = V{x[d[0]], x[d[1]]} - ce;
            V e = V{x[d[2]], x[d[3]]} - ce;
            V p = V{x[d[6]], x[d[7]]} - ce;
            double r = magnitude(s), r_e = magnitude(e), r_p = magnitude(p);
            if (r < EPSILON || r_e < EPSILON || r_p < EPSILON) {
                *degenerate = true;
                return;
            }
            V u_s = s * (1.0 / r);
            V u_e = e * (1.0 / r_e);
            V e_proj = e * (r / r_e);
            double j_s[2][2], j_e[2][2], j_p[2][2];
            switch (classify_point_arc_coincident(s, e_proj, p)) {
                case PAC_INTERIOR: {
                    V u_p = p * (1.0 / r_p);
                    double r_over_rp = r / r_p;
                    j_s[0][0] = u_p.x * u_s.x; j_s[0][1] = u_p.y * u_s.x;
                    j_s[1][0] = u_p.x * u_s.y; j_s[1][1] = u_p.y * u_s.y;
                    j_e[0][0] = j_e[0][1] = j_e[1][0] = j_e[1][1] = 0.0;
                    j_p[0][0] = (r_over_rp - 1.0) - r_over_rp * u_p.x * u_p.x;
                    j_p[0][1] = -r_over_rp * u_p.y * u_p.x;
                    j_p[1][0] = -r_over_rp * u_p.x * u_p.y;
                    j_p[1][1] = (r_over_rp - 1.0) - r_over_rp * u_p.y * u_p.y;
                } break;
                case PAC_END: {
                    double r_over_re = r / r_e;
                    j_s[0][0] = u_e.x * u_s.x; j_s[0][1] = u_e.y * u_s.x;
                    j_s[1][0] = u_e.x * u_s.y; j_s[1][1] = u_e.y * u_s.y;
                    j_e[0][0] = r_over_re * (1.0 - u_e.x * u_e.x);
                    j_e[0][1] = -r_over_re * u_e.y * u_e.x;
                    j_e[1][0] = -r_over_re * u_e.x * u_e.y;
                    j_e[1][1] = r_over_re * (1.0 - u_e.y * u_e.y);
                    j_p[0][0] = -1.0; j_p[0][1] = 0.0; j_p[1][0] = 0.0; j_p[1][1] = -1.0;
                } break;
                default: {
                    j_s[0][0] = 1.0; j_s[0][1] = 0.0; j_s[1][0] = 0.0; j_s[1][1] = 1.0;
                    j_e[0][0] = j_e[0][1] = j_e[1][0] = j_e[1][1] = 0.0;
                    j_p[0][0] = -1.0; j_p[0][1] = 0.0; j_p[1][0] = 0.0; j_p[1][1] = -1.0;
                } break;
            }
            double j_o[2][2] = {{-(j_s[0][0] + j_e[0][0] + j_p[0][0]), -(j_s[0][1] + j_e[0][1] + j_p[0][1])},
                                {-(j_s[1][0] + j_e[1][0] + j_p[1][0]), -(j_s[1][1] + j_e[1][1] + j_p[1][1])}};
            // ids: c, s, e, p (:1995-2062)
            const uint32_t idv[8] = {d[4], d[5], d[0], d[1], d[2], d[3], d[6], d[7]};
            const double v0[8] = {j_o[0][0], j_o[1][0], j_s[0][0], j_s[1][0], j_e[0][0], j_e[1][0], j_p[0][0], j_p[1][0]};
            const double v1[8] = {j_o[0][1], j_o[1][1], j_s[0][1], j_s[1][1], j_e[0][1], j_e[1][1], j_p[0][1], j_p[1][1]};
            for (int k = 0; k < 8; ++k) {
                row0.push_back({idv[k], v0[k]});
                row1.push_back({idv[k], v1[k]});
            }
        } break;
        case K_ARC_LENGTH: {
            double ax = x[d[0]], ay = x[d[1]], cx = x[d[4]], cy = x[d[5]];
            double ux = ax - cx, uy = ay - cy;
            double r2 = ux * ux + uy * uy;
            if (r2 <= EPSILON * EPSILON) {
                *degenerate = true;
                return;
            }
            double r = std::sqrt(r2);
            double alpha = c.p0 / r;
            double sa = orc_sin(alpha), ca = orc_cos(alpha);
            double rux = ca * ux - sa * uy;
            double ruy = sa * ux + ca * uy;
            double k = c.p0 / (r2 * r);
            const double v0[6] = {-ca - ruy * ux * k, sa - ruy * uy * k, 1.0, 0.0, -1.0 + ca + ruy * ux * k,
                                  -sa + ruy * uy * k};
            const double v1[6] = {-sa + rux * ux * k, -ca + rux * uy * k, 0.0, 1.0, sa - rux * ux * k,
                                  -1.0 + ca - rux * uy * k};
            for (int q = 0; q < 6; ++q) {
                row0.push_back({d[q], v0[q]});
                row1.push_back({d[q], v1[q]});
            }
        } break;
        case K_ARC_ANGLE: {
            const uint32_t idv[8] = {d[4], d[5], d[0], d[1], d[4], d[5], d[2], d[3]};
            lines_at_angle_jacobian(idv, x[d[4]], x[d[5]], x[d[0]], x[d[1]], x[d[4]], x[d[5]], x[d[2]], x[d[3]],
                                    Rot{c.p0, c.p1}, row0, degenerate);
        } break;
        case K_POINTS_AT_ANGLE: {
            V p0{x[d[0]], x[d[1]]}, p1{x[d[2]], x[d[3]]}, p2{x[d[4]], x[d[5]]};
            V u = p1 - p0, v = p2 - p0;
            double len_u = magnitude(u), len_v = magnitude(v);
            if (len_u <= EPSILON || len_v <= EPSILON) {
                *degenerate = true;
                return;
            }
            double inv_len_u = 1.0 / len_u, inv_len_v = 1.0 / len_v;
            V u_hat = u * inv_len_u, v_hat = v * inv_len_v;
            Rot rot{c.p0, c.p1};
            double s = (len_u + len_v) * 0.5;
            V rot_e1 = apply(rot, V{1.0, 0.0});
            V rot_e2 = apply(rot, V{0.0, 1.0});
            double inv_s = 1.0 / s;
            V rot_u = apply(rot, u);
            V res = (v * len_u - rot_u * len_v) * inv_s;
            V half_res = res * 0.5;
            V dr_du0 = ((v - half_res) * u_hat.x - rot_e1 * len_v) * inv_s;
            V dr_du1 = ((v - half_res) * u_hat.y - rot_e2 * len_v) * inv_s;
            V dr_dv0 = (V{len_u, 0.0} - (rot_u + half_res) * v_hat.x) * inv_s;
            V dr_dv1 = (V{0.0, len_u} - (rot_u + half_res) * v_hat.y) * inv_s;
            row0.push_back({d[0], -(dr_du0.x + dr_dv0.x)});
            row0.push_back({d[1], -(dr_du1.x + dr_dv1.x)});
            row0.push_back({d[2], dr_du0.x});
            row0.push_back({d[3], dr_du1.x});
            row0.push_back({d[4], dr_dv0.x});
            row0.push_back({d[5], dr_dv1.x});
            row1.push_back({d[0], -(dr_du0.y + dr_dv0.y)});
            row1.push_back({d[1], -(dr_du1.y + dr_dv1.y)});
            row1.push_back({d[2], dr_du0.y});
            row1.push_back({d[3], dr_du1.y});
            row1.push_back({d[4], dr_dv0.y});
            row1.push_back({d[5], dr_dv1.y});
        } break;
        default: break;
    }
}

}  // namespace orc
