// ORACLE — TEST INFRASTRUCTURE ONLY. Nothing under ezpz_b200/ may include, link or call this.
//
// CPU restatement of the scalar math the reference gets from the un-vendored crate
// `libm` 0.2.16 (Cargo.lock:1139-1140; a Rust port of musl libc's libm, itself FreeBSD msun).
// Reference call sites: ezpz/src/vector.rs:16 (hypot), :21 (pow(.,2)), :73 (atan2), :118 (sincos);
// ezpz/src/constraints.rs:581,692-693,728,887-888,1557-1558,1698-1701,1756-1758,2089-2090,2379,
// 2455,2462; ezpz/src/solver/newton.rs:53,108 (fmax).
//
// What is restated, and how faithfully:
//   orc_hypot  — musl src/math/hypot.c: exponent compare, 2^±700 rescale, Dekker split squares,
//                one sqrt.  Built only from IEEE +,-,*,sqrt, so it is bit-reproducible anywhere.
//   orc_sin/orc_cos — musl __sin.c/__cos.c kernels and the "medium" branch of __rem_pio2.c
//                (3-stage Cody–Waite with pio2_1/2/3).  The Payne–Hanek branch for |x| >= 2^20*pi/2
//                is NOT restated: beyond that bound the medium reduction is applied anyway and the
//                result loses accuracy (documented deviation; an arc whose length/radius ratio
//                exceeds 1.6e6 rad is not a sketch).
//   orc_atan2  — musl atan2.c + atan.c.
//   pow        — libm::pow special-cases y == 2 as x*x exactly (FreeBSD e_pow.c "y is 2"), which is
//                what every pow(.,2.0) call site hits.  pow(t,-2.0) and pow(t,1.5) go through the
//                general log/exp path in libm (error < 1 ulp, not correctly rounded); they are
//                restated here as 1/(t*t) and t*sqrt(t), which can differ from libm in the last ulp.
//                Only the Jacobians of PointLineDistance / Vertical- / HorizontalPointLineDistance
//                use them (constraints.rs:1698-1701,1756-1758,2462).
//   orc_fmax   — NaN-ignoring maximum (libm::fmax).
//
// The same algorithms are written a second time, independently, for the device in
// ezpz_b200/csrc/dmath.cuh; tests/ compare the two bit-for-bit.
//
// Compile with -ffp-contract=off: every '*' followed by '+' below is two roundings, as in Rust.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

static inline uint64_t bits_of(double x) { uint64_t u; std::memcpy(&u, &x, 8); return u; }
static inline double from_bits(uint64_t u) { double x; std::memcpy(&x, &u, 8); return x; }
static inline uint32_t hi_word(double x) { return (uint32_t)(bits_of(x) >> 32); }
static inline uint32_t lo_word(double x) { return (uint32_t)(bits_of(x)); }

// ---------------------------------------------------------------- hypot (musl hypot.c)
static inline void sq_split(double* hi, double* lo, double x) {
    const double SPLIT = 134217729.0;  // 0x1p27 + 1
    double xc = x * SPLIT;
    double xh = x - xc + xc;
    double xl = x - xh;
    *hi = x * x;
    *lo = xh * xh - *hi + 2 * xh * xl + xl * xl;
}

static inline double orc_hypot(double x, double y) {
    uint64_t ux = bits_of(x) & (~0ULL >> 1);
    uint64_t uy = bits_of(y) & (~0ULL >> 1);
    if (ux < uy) { uint64_t t = ux; ux = uy; uy = t; }
    int ex = (int)(ux >> 52);
    int ey = (int)(uy >> 52);
    x = from_bits(ux);
    y = from_bits(uy);
    // hypot(inf, nan) == inf
    if (ey == 0x7ff) return y;
    if (ex == 0x7ff || uy == 0) return x;
    if (ex - ey > 64) return x + y;
    double z = 1.0;
    if (ex > 0x3ff + 510) {
        z = 0x1p700; x *= 0x1p-700; y *= 0x1p-700;
    } else if (ey < 0x3ff - 450) {
        z = 0x1p-700; x *= 0x1p700; y *= 0x1p700;
    }
    double hx, lx, hy, ly;
    sq_split(&hx, &lx, x);
    sq_split(&hy, &ly, y);
    return z * std::sqrt(ly + lx + hy + hx);
}

// ---------------------------------------------------------------- sin / cos kernels (musl __sin.c, __cos.c)
static inline double k_sin(double x, double y, int iy) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = x * x;
    double w = z * z;
    double r = S2 + z * (S3 + z * S4) + z * w * (S5 + z * S6);
    double v = z * x;
    if (iy == 0) return x + v * (S1 + z * r);
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

static inline double k_cos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = x * x;
    double w = z * z;
    double r = z * (C1 + z * (C2 + z * C3)) + w * w * (C4 + z * (C5 + z * C6));
    double hz = 0.5 * z;
    w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}

// musl __rem_pio2.c, "medium" branch only (see header note).  Returns n mod 4 information in n,
// remainder in y[0] + y[1].
static inline int rem_pio2_medium(double x, double* y) {
    const double toint = 1.5 / 2.220446049250313e-16;  // 1.5/EPS
    const double invpio2 = 6.36619772367581382433e-01;
    const double pio2_1 = 1.57079632673412561417e+00, pio2_1t = 6.07710050650619224932e-11;
    const double pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21;
    const double pio2_3 = 2.02226624871116645580e-21, pio2_3t = 8.47842766036889956997e-32;
    double fn = x * invpio2 + toint - toint;
    int n = (int)fn;
    double r = x - fn * pio2_1;
    double w = fn * pio2_1t;
    int ex = (int)((hi_word(x) >> 20) & 0x7ff);
    y[0] = r - w;
    int ey = (int)((hi_word(y[0]) >> 20) & 0x7ff);
    if (ex - ey > 16) {
        double t = r;
        w = fn * pio2_2;
        r = t - w;
        w = fn * pio2_2t - ((t - r) - w);
        y[0] = r - w;
        ey = (int)((hi_word(y[0]) >> 20) & 0x7ff);
        if (ex - ey > 49) {
            t = r;
            w = fn * pio2_3;
            r = t - w;
            w = fn * pio2_3t - ((t - r) - w);
            y[0] = r - w;
        }
    }
    y[1] = (r - y[0]) - w;
    return n;
}

static inline double orc_sin(double x) {
    uint32_t ix = hi_word(x) & 0x7fffffff;
    if (ix <= 0x3fe921fb) {             // |x| ~< pi/4
        if (ix < 0x3e500000) return x;  // |x| < 2**-26
        return k_sin(x, 0.0, 0);
    }
    if (ix >= 0x7ff00000) return x - x;
    double y[2];
    int n = rem_pio2_medium(x, y);
    switch (n & 3) {
        case 0: return k_sin(y[0], y[1], 1);
        case 1: return k_cos(y[0], y[1]);
        case 2: return -k_sin(y[0], y[1], 1);
        default: return -k_cos(y[0], y[1]);
    }
}

static inline double orc_cos(double x) {
    uint32_t ix = hi_word(x) & 0x7fffffff;
    if (ix <= 0x3fe921fb) {
        if (ix < 0x3e46a09e) return 1.0;  // |x| < 2**-27 * sqrt(2)
        return k_cos(x, 0.0);
    }
    if (ix >= 0x7ff00000) return x - x;
    double y[2];
    int n = rem_pio2_medium(x, y);
    switch (n & 3) {
        case 0: return k_cos(y[0], y[1]);
        case 1: return -k_sin(y[0], y[1], 1);
        case 2: return -k_cos(y[0], y[1]);
        default: return k_sin(y[0], y[1], 1);
    }
}

// ---------------------------------------------------------------- atan / atan2 (musl atan.c, atan2.c)
static inline double orc_atan(double x) {
    static const double atanhi[4] = {4.63647609000806093515e-01, 7.85398163397448278999e-01,
                                     9.82793723247329054082e-01, 1.57079632679489655800e+00};
    static const double atanlo[4] = {2.26987774529616870924e-17, 3.06161699786838301793e-17,
                                     1.39033110312309984516e-17, 6.12323399573676603587e-17};
    static const double aT[11] = {3.33333333333329318027e-01,  -1.99999999998764832476e-01,
                                  1.42857142725034663711e-01,  -1.11111104054623557880e-01,
                                  9.09088713343650656196e-02,  -7.69187620504482999495e-02,
                                  6.66107313738753120669e-02,  -5.83357013379057348645e-02,
                                  4.97687799461593236017e-02,  -3.65315727442169155270e-02,
                                  1.62858201153657823623e-02};
    uint32_t ix = hi_word(x);
    uint32_t sign = ix >> 31;
    ix &= 0x7fffffff;
    int id;
    if (ix >= 0x44100000) {  // |x| >= 2^66
        if (x != x) return x;
        double z = atanhi[3] + 0x1p-120;
        return sign ? -z : z;
    }
    if (ix < 0x3fdc0000) {      // |x| < 0.4375
        if (ix < 0x3e400000) {  // |x| < 2^-27
            return x;
        }
        id = -1;
    } else {
        x = std::fabs(x);
        if (ix < 0x3ff30000) {      // |x| < 1.1875
            if (ix < 0x3fe60000) {  // 7/16 <= |x| < 11/16
                id = 0;
                x = (2.0 * x - 1.0) / (2.0 + x);
            } else {                // 11/16 <= |x| < 19/16
                id = 1;
                x = (x - 1.0) / (x + 1.0);
            }
        } else {
            if (ix < 0x40038000) {  // |x| < 2.4375
                id = 2;
                x = (x - 1.5) / (1.0 + 1.5 * x);
            } else {                // 2.4375 <= |x| < 2^66
                id = 3;
                x = -1.0 / x;
            }
        }
    }
    double z = x * x;
    double w = z * z;
    double s1 = z * (aT[0] + w * (aT[2] + w * (aT[4] + w * (aT[6] + w * (aT[8] + w * aT[10])))));
    double s2 = w * (aT[1] + w * (aT[3] + w * (aT[5] + w * (aT[7] + w * aT[9]))));
    if (id < 0) return x - x * (s1 + s2);
    z = atanhi[id] - (x * (s1 + s2) - atanlo[id] - x);
    return sign ? -z : z;
}

static inline double orc_atan2(double y, double x) {
    const double pi = 3.1415926535897931160E+00, pi_lo = 1.2246467991473531772E-16;
    if (x != x || y != y) return x + y;
    uint32_t ix = hi_word(x), iy = hi_word(y);
    uint32_t lx = lo_word(x), ly = lo_word(y);
    if (((ix - 0x3ff00000) | lx) == 0) return orc_atan(y);  // x = 1.0
    uint32_t m = ((iy >> 31) & 1) | ((ix >> 30) & 2);        // 2*sign(x) + sign(y)
    ix &= 0x7fffffff;
    iy &= 0x7fffffff;
    if ((iy | ly) == 0) {  // y = 0
        switch (m) {
            case 0:
            case 1: return y;
            case 2: return pi;
            default: return -pi;
        }
    }
    if ((ix | lx) == 0) return (m & 1) ? -pi / 2 : pi / 2;  // x = 0
    if (ix == 0x7ff00000) {                                  // x = INF
        if (iy == 0x7ff00000) {
            switch (m) {
                case 0: return pi / 4;
                case 1: return -pi / 4;
                case 2: return 3 * pi / 4;
                default: return -3 * pi / 4;
            }
        } else {
            switch (m) {
                case 0: return 0.0;
                case 1: return -0.0;
                case 2: return pi;
                default: return -pi;
            }
        }
    }
    // |y/x| > 0x1p64
    if (ix + (64 << 20) < iy || iy == 0x7ff00000) return (m & 1) ? -pi / 2 : pi / 2;
    double z;
    // z = atan(|y/x|) without spurious underflow
    if ((m & 2) && iy + (64 << 20) < ix)  // |y/x| < 0x1p-64, x<0
        z = 0;
    else
        z = orc_atan(std::fabs(y / x));
    switch (m) {
        case 0: return z;
        case 1: return -z;
        case 2: return pi - (z - pi_lo);
        default: return (z - pi_lo) - pi;
    }
}

// ---------------------------------------------------------------- small helpers with Rust/libm semantics
static inline double orc_fmax(double a, double b) {  // libm::fmax: NaN-ignoring
    if (a != a) return b;
    if (b != b) return a;
    return a < b ? b : a;
}
static inline double orc_pow2(double x) { return x * x; }                    // libm::pow(x, 2.0)
static inline double orc_pow_m2(double x) { return 1.0 / (x * x); }          // libm::pow(x, -2.0) (see note)
static inline double orc_pow_1p5(double x) { return x * std::sqrt(x); }      // libm::pow(x, 1.5)  (see note)
static inline double orc_signum(double x) {                                  // f64::signum
    if (x != x) return x;
    return std::signbit(x) ? -1.0 : 1.0;
}
static inline double orc_rem_euclid(double a, double b) {                    // f64::rem_euclid
    double r = std::fmod(a, b);
    return r < 0.0 ? r + std::fabs(b) : r;
}

}  // namespace orc
