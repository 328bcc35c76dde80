// dmath.cuh — binary64 scalar math with the bits of the `libm` crate 0.2.16 (musl/FreeBSD msun),
// usable from device kernels and from host code.
//
// The reference calls libm for hypot / sin / cos / atan2 / pow / fmax (ezpz/src/vector.rs:16,21,73,118;
// ezpz/src/constraints.rs:581,692-693,728,887-888,2089-2090,2379,2455,2462; clippy.toml bans std math).
// CUDA's own hypot()/sin()/cos()/atan2() are 1-2 ulp functions with different last bits, and every
// accept/reject decision of the LM loop is a strict floating-point comparison (newton.rs:118), so the
// solve path carries its own implementations, built only from IEEE add/mul/div/sqrt (which CUDA rounds
// correctly in binary64) and integer bit tests:
//   ez_hypot : exponent test, 2^+-700 rescale, Dekker-split exact squares, one sqrt (musl hypot.c)
//   ez_sin, ez_cos : msun kernels on [-pi/4, pi/4] + three-constant Cody-Waite reduction
//              ("medium" branch of __rem_pio2.c; arguments beyond 2^20*pi/2 are reduced the same way
//              and lose accuracy - an arc subtending more than 1.6e6 rad is not a sketch)
//   ez_atan2 : msun atan2.c / atan.c
//   ez_pow2(x) = x*x (libm::pow special-cases y == 2); ez_pow_m2(x) = 1/(x*x); ez_pow_1p5(x) = x*sqrt(x)
//   ez_fmax  : NaN-ignoring maximum
// All products/sums are written through EZ_MUL/EZ_ADD/EZ_SUB so that no FMA contraction can change a
// rounding, whatever -fmad says.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define EZ_HD __host__ __device__ __forceinline__
#else
#define EZ_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define EZ_MUL(a, b) __dmul_rn((a), (b))
#define EZ_ADD(a, b) __dadd_rn((a), (b))
#define EZ_SUB(a, b) __dsub_rn((a), (b))
#define EZ_DIV(a, b) __ddiv_rn((a), (b))
#define EZ_SQRT(a) __dsqrt_rn((a))
#define EZ_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
// Host translation units are compiled with -ffp-contract=off (see __graft_entry__.build()).
#define EZ_MUL(a, b) ((a) * (b))
#define EZ_ADD(a, b) ((a) + (b))
#define EZ_SUB(a, b) ((a) - (b))
#define EZ_DIV(a, b) ((a) / (b))
#define EZ_SQRT(a) std::sqrt((a))
#define EZ_FMA(a, b, c) std::fma((a), (b), (c))
#endif

namespace ezm {

EZ_HD uint64_t d2u(double x) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    std::memcpy(&u, &x, sizeof u);
    return u;
#endif
}
EZ_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    std::memcpy(&x, &u, sizeof x);
    return x;
#endif
}
EZ_HD uint32_t hiword(double x) { return (uint32_t)(d2u(x) >> 32); }
EZ_HD uint32_t loword(double x) { return (uint32_t)d2u(x); }
EZ_HD double ez_abs(double x) { return u2d(d2u(x) & 0x7fffffffffffffffULL); }
EZ_HD bool ez_isnan(double x) { return x != x; }
EZ_HD bool ez_isfinite(double x) { return ((d2u(x) >> 52) & 0x7ff) != 0x7ff; }

// NaN-ignoring max (libm::fmax; newton.rs:53,108).
EZ_HD double ez_fmax(double a, double b) {
    if (ez_isnan(a)) return b;
    if (ez_isnan(b)) return a;
    return a < b ? b : a;
}
// f64::signum: +-1 by sign bit, NaN stays NaN (constraints.rs:1057,1121-1125).
EZ_HD double ez_signum(double x) {
    if (ez_isnan(x)) return x;
    return (d2u(x) >> 63) ? -1.0 : 1.0;
}

// Exact square as hi + lo (Dekker / Veltkamp split with 2^27 + 1).
EZ_HD void ez_sq(double x, double& hi, double& lo) {
    const double kSplit = 134217729.0;
    double xc = EZ_MUL(x, kSplit);
    double xh = EZ_ADD(EZ_SUB(x, xc), xc);
    double xl = EZ_SUB(x, xh);
    hi = EZ_MUL(x, x);
    // xh*xh - hi + 2*xh*xl + xl*xl, left to right
    double t = EZ_SUB(EZ_MUL(xh, xh), hi);
    t = EZ_ADD(t, EZ_MUL(EZ_MUL(2.0, xh), xl));
    lo = EZ_ADD(t, EZ_MUL(xl, xl));
}

EZ_HD double ez_hypot(double x, double y) {
    uint64_t ax = d2u(x) & 0x7fffffffffffffffULL;
    uint64_t ay = d2u(y) & 0x7fffffffffffffffULL;
    if (ax < ay) {
        uint64_t t = ax;
        ax = ay;
        ay = t;
    }
    const int ex = (int)(ax >> 52), ey = (int)(ay >> 52);
    double big = u2d(ax), small = u2d(ay);
    if (ey == 0x7ff) return small;              // hypot(inf, nan) == inf; nan otherwise
    if (ex == 0x7ff || ay == 0) return big;
    if (ex - ey > 64) return EZ_ADD(big, small);
    double scale = 1.0;
    if (ex > 0x3ff + 510) {
        scale = 0x1p700;
        big = EZ_MUL(big, 0x1p-700);
        small = EZ_MUL(small, 0x1p-700);
    } else if (ey < 0x3ff - 450) {
        scale = 0x1p-700;
        big = EZ_MUL(big, 0x1p700);
        small = EZ_MUL(small, 0x1p700);
    }
    double hb, lb, hs, ls;
    ez_sq(big, hb, lb);
    ez_sq(small, hs, ls);
    // ly + lx + hy + hx with x the larger
    double s = EZ_ADD(EZ_ADD(EZ_ADD(ls, lb), hs), hb);
    return EZ_MUL(scale, EZ_SQRT(s));
}

EZ_HD double ez_pow2(double x) { return EZ_MUL(x, x); }
EZ_HD double ez_pow_m2(double x) { return EZ_DIV(1.0, EZ_MUL(x, x)); }
EZ_HD double ez_pow_1p5(double x) { return EZ_MUL(x, EZ_SQRT(x)); }

// ---- sin / cos ------------------------------------------------------------------------------
// Minimax kernels on |x| <= pi/4 with a tail y (x + y is the reduced argument).
EZ_HD double ez_ksin(double x, double y, bool have_tail) {
    const double s1 = -1.66666666666666324348e-01, s2 = 8.33333333332248946124e-03,
                 s3 = -1.98412698298579493134e-04, s4 = 2.75573137070700676789e-06,
                 s5 = -2.50507602534068634195e-08, s6 = 1.58969099521155010221e-10;
    double z = EZ_MUL(x, x);
    double w = EZ_MUL(z, z);
    // r = s2 + z*(s3 + z*s4) + z*w*(s5 + z*s6)
    double r = EZ_ADD(EZ_ADD(s2, EZ_MUL(z, EZ_ADD(s3, EZ_MUL(z, s4)))),
                      EZ_MUL(EZ_MUL(z, w), EZ_ADD(s5, EZ_MUL(z, s6))));
    double v = EZ_MUL(z, x);
    if (!have_tail) return EZ_ADD(x, EZ_MUL(v, EZ_ADD(s1, EZ_MUL(z, r))));
    // x - ((z*(0.5*y - v*r) - y) - v*s1)
    double t = EZ_SUB(EZ_MUL(z, EZ_SUB(EZ_MUL(0.5, y), EZ_MUL(v, r))), y);
    return EZ_SUB(x, EZ_SUB(t, EZ_MUL(v, s1)));
}

EZ_HD double ez_kcos(double x, double y) {
    const double c1 = 4.16666666666666019037e-02, c2 = -1.38888888888741095749e-03,
                 c3 = 2.48015872894767294178e-05, c4 = -2.75573143513906633035e-07,
                 c5 = 2.08757232129817482790e-09, c6 = -1.13596475577881948265e-11;
    double z = EZ_MUL(x, x);
    double w = EZ_MUL(z, z);
    // r = z*(c1 + z*(c2 + z*c3)) + w*w*(c4 + z*(c5 + z*c6))
    double ra = EZ_MUL(z, EZ_ADD(c1, EZ_MUL(z, EZ_ADD(c2, EZ_MUL(z, c3)))));
    double rb = EZ_MUL(EZ_MUL(w, w), EZ_ADD(c4, EZ_MUL(z, EZ_ADD(c5, EZ_MUL(z, c6)))));
    double r = EZ_ADD(ra, rb);
    double hz = EZ_MUL(0.5, z);
    double one_m = EZ_SUB(1.0, hz);
    // w + (((1 - w) - hz) + (z*r - x*y))
    double corr = EZ_ADD(EZ_SUB(EZ_SUB(1.0, one_m), hz), EZ_SUB(EZ_MUL(z, r), EZ_MUL(x, y)));
    return EZ_ADD(one_m, corr);
}

// x = n*(pi/2) + (y0 + y1), |y0 + y1| <= pi/4 (+ a hair); returns n.
EZ_HD int ez_reduce_pio2(double x, double& y0, double& y1) {
    const double to_int = 6755399441055744.0;  // 1.5 / DBL_EPSILON
    const double inv_pio2 = 6.36619772367581382433e-01;
    const double p1 = 1.57079632673412561417e+00, p1t = 6.07710050650619224932e-11;
    const double p2 = 6.07710050630396597660e-11, p2t = 2.02226624879595063154e-21;
    const double p3 = 2.02226624871116645580e-21, p3t = 8.47842766036889956997e-32;
    double fn = EZ_SUB(EZ_ADD(EZ_MUL(x, inv_pio2), to_int), to_int);
    int n = (int)fn;
    double r = EZ_SUB(x, EZ_MUL(fn, p1));
    double w = EZ_MUL(fn, p1t);
    const int ex = (int)((hiword(x) >> 20) & 0x7ff);
    y0 = EZ_SUB(r, w);
    int ey = (int)((hiword(y0) >> 20) & 0x7ff);
    if (ex - ey > 16) {
        double t = r;
        w = EZ_MUL(fn, p2);
        r = EZ_SUB(t, w);
        w = EZ_SUB(EZ_MUL(fn, p2t), EZ_SUB(EZ_SUB(t, r), w));
        y0 = EZ_SUB(r, w);
        ey = (int)((hiword(y0) >> 20) & 0x7ff);
        if (ex - ey > 49) {
            t = r;
            w = EZ_MUL(fn, p3);
            r = EZ_SUB(t, w);
            w = EZ_SUB(EZ_MUL(fn, p3t), EZ_SUB(EZ_SUB(t, r), w));
            y0 = EZ_SUB(r, w);
        }
    }
    y1 = EZ_SUB(EZ_SUB(r, y0), w);
    return n;
}

EZ_HD void ez_sincos(double x, double& s, double& c) {
    const uint32_t ix = hiword(x) & 0x7fffffffu;
    if (ix <= 0x3fe921fbu) {
        s = (ix < 0x3e500000u) ? x : ez_ksin(x, 0.0, false);
        c = (ix < 0x3e46a09eu) ? 1.0 : ez_kcos(x, 0.0);
        return;
    }
    if (ix >= 0x7ff00000u) {
        s = c = EZ_SUB(x, x);
        return;
    }
    double y0, y1;
    const int n = ez_reduce_pio2(x, y0, y1);
    const double ks = ez_ksin(y0, y1, true);
    const double kc = ez_kcos(y0, y1);
    switch (n & 3) {
        case 0: s = ks; c = kc; break;
        case 1: s = kc; c = -ks; break;
        case 2: s = -ks; c = -kc; break;
        default: s = -kc; c = ks; break;
    }
}
EZ_HD double ez_sin(double x) {
    double s, c;
    ez_sincos(x, s, c);
    return s;
}
EZ_HD double ez_cos(double x) {
    double s, c;
    ez_sincos(x, s, c);
    return c;
}

// ---- atan / atan2 ---------------------------------------------------------------------------
EZ_HD double ez_atan(double x) {
    const double hi0 = 4.63647609000806093515e-01, hi1 = 7.85398163397448278999e-01,
                 hi2 = 9.82793723247329054082e-01, hi3 = 1.57079632679489655800e+00;
    const double lo0 = 2.26987774529616870924e-17, lo1 = 3.06161699786838301793e-17,
                 lo2 = 1.39033110312309984516e-17, lo3 = 6.12323399573676603587e-17;
    const double a0 = 3.33333333333329318027e-01, a1 = -1.99999999998764832476e-01,
                 a2 = 1.42857142725034663711e-01, a3 = -1.11111104054623557880e-01,
                 a4 = 9.09088713343650656196e-02, a5 = -7.69187620504482999495e-02,
                 a6 = 6.66107313738753120669e-02, a7 = -5.83357013379057348645e-02,
                 a8 = 4.97687799461593236017e-02, a9 = -3.65315727442169155270e-02,
                 a10 = 1.62858201153657823623e-02;
    uint32_t ix = hiword(x);
    const bool neg = (ix >> 31) != 0;
    ix &= 0x7fffffffu;
    if (ix >= 0x44100000u) {  // |x| >= 2^66 or NaN
        if (ez_isnan(x)) return x;
        double z = EZ_ADD(hi3, 0x1p-120);
        return neg ? -z : z;
    }
    int seg = -1;
    double hi = 0.0, lo = 0.0;
    if (ix < 0x3fdc0000u) {  // |x| < 7/16
        if (ix < 0x3e400000u) return x;
    } else {
        x = ez_abs(x);
        if (ix < 0x3ff30000u) {
            if (ix < 0x3fe60000u) {
                seg = 0; hi = hi0; lo = lo0;
                x = EZ_DIV(EZ_SUB(EZ_MUL(2.0, x), 1.0), EZ_ADD(2.0, x));
            } else {
                seg = 1; hi = hi1; lo = lo1;
                x = EZ_DIV(EZ_SUB(x, 1.0), EZ_ADD(x, 1.0));
            }
        } else if (ix < 0x40038000u) {
            seg = 2; hi = hi2; lo = lo2;
            x = EZ_DIV(EZ_SUB(x, 1.5), EZ_ADD(1.0, EZ_MUL(1.5, x)));
        } else {
            seg = 3; hi = hi3; lo = lo3;
            x = EZ_DIV(-1.0, x);
        }
    }
    double z = EZ_MUL(x, x);
    double w = EZ_MUL(z, z);
    // odd and even halves of the polynomial, Horner in w
    double e = EZ_ADD(a8, EZ_MUL(w, a10));
    e = EZ_ADD(a6, EZ_MUL(w, e));
    e = EZ_ADD(a4, EZ_MUL(w, e));
    e = EZ_ADD(a2, EZ_MUL(w, e));
    e = EZ_ADD(a0, EZ_MUL(w, e));
    double s1 = EZ_MUL(z, e);
    double o = EZ_ADD(a7, EZ_MUL(w, a9));
    o = EZ_ADD(a5, EZ_MUL(w, o));
    o = EZ_ADD(a3, EZ_MUL(w, o));
    o = EZ_ADD(a1, EZ_MUL(w, o));
    double s2 = EZ_MUL(w, o);
    double xs = EZ_MUL(x, EZ_ADD(s1, s2));
    if (seg < 0) return EZ_SUB(x, xs);
    z = EZ_SUB(hi, EZ_SUB(EZ_SUB(xs, lo), x));
    return neg ? -z : z;
}

EZ_HD double ez_atan2(double y, double x) {
    const double pi = 3.1415926535897931160E+00, pi_lo = 1.2246467991473531772E-16;
    if (ez_isnan(x) || ez_isnan(y)) return EZ_ADD(x, y);
    uint32_t ix = hiword(x), iy = hiword(y);
    const uint32_t lx = loword(x), ly = loword(y);
    if (((ix - 0x3ff00000u) | lx) == 0) return ez_atan(y);
    const uint32_t quad = ((iy >> 31) & 1u) | ((ix >> 30) & 2u);
    ix &= 0x7fffffffu;
    iy &= 0x7fffffffu;
    if ((iy | ly) == 0) {
        if (quad < 2) return y;
        return quad == 2 ? pi : -pi;
    }
    if ((ix | lx) == 0) return (quad & 1u) ? -EZ_DIV(pi, 2.0) : EZ_DIV(pi, 2.0);
    if (ix == 0x7ff00000u) {
        if (iy == 0x7ff00000u) {
            switch (quad) {
                case 0: return EZ_DIV(pi, 4.0);
                case 1: return -EZ_DIV(pi, 4.0);
                case 2: return EZ_DIV(EZ_MUL(3.0, pi), 4.0);
                default: return -EZ_DIV(EZ_MUL(3.0, pi), 4.0);
            }
        }
        switch (quad) {
            case 0: return 0.0;
            case 1: return -0.0;
            case 2: return pi;
            default: return -pi;
        }
    }
    if (ix + (64u << 20) < iy || iy == 0x7ff00000u) return (quad & 1u) ? -EZ_DIV(pi, 2.0) : EZ_DIV(pi, 2.0);
    double z;
    if ((quad & 2u) && iy + (64u << 20) < ix) z = 0.0;
    else z = ez_atan(ez_abs(EZ_DIV(y, x)));
    switch (quad) {
        case 0: return z;
        case 1: return -z;
        case 2: return EZ_SUB(pi, EZ_SUB(z, pi_lo));
        default: return EZ_SUB(EZ_SUB(z, pi_lo), pi);
    }
}

// f64::rem_euclid(a, 2*pi) for a in [-pi, pi] (the only use, constraints.rs:2596-2597): fmod is the
// identity there, so only the sign fix-up remains.
EZ_HD double ez_wrap_0_2pi(double a) {
    const double two_pi = 6.283185307179586;
    return a < 0.0 ? EZ_ADD(a, two_pi) : a;
}

}  // namespace ezm
