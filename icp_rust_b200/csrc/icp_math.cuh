// icp_math.cuh — scalar f64 device maths of the ICP update step.
//
// Every function restates one reference function (file:line relative to tier4/icp_rust) as a
// fixed sequence of IEEE-754 binary64 operations in source order.  The translation unit is
// compiled with -fmad=false, so nvcc never contracts a*b+c into an FMA (rustc does not
// either); double-precision '/' and sqrt() are correctly rounded on sm_100a.  Same inputs
// therefore give the same bits as the reference's arithmetic.
#pragma once
#include "simt_compat.h"

namespace icp {

constexpr double kHuberK = 1.345;               // src/lib.rs:32
constexpr double kPpf34 = 1.482602218505602;    // src/stats.rs:42
constexpr double kDeltaNormThreshold = 1e-6;    // src/lib.rs:60 (compared with the SQUARED norm)
constexpr int kInnerMaxIter = 200;              // src/lib.rs:61
constexpr double kF64Max = 1.7976931348623157e308;  // f64::MAX, src/lib.rs:63

enum BreakReason : uint32_t { kBreakCap = 0, kBreakNone = 1, kBreakDelta = 2, kBreakError = 3 };

// SE(2) transform as the reference stores it: full 2x2 rotation + translation
// (src/transform.rs:6-10); never re-orthonormalised.
struct Transform {
    double r00, r01, r10, r11, tx, ty;
};

ICP_DEV double bits_to_f64(uint64_t b) { return __longlong_as_double((long long)b); }
ICP_DEV uint64_t f64_to_bits(double d) { return (uint64_t)__double_as_longlong(d); }
#define ICP_HEXD(hi, lo) ((((uint64_t)(hi)) << 32) | (uint64_t)(lo))

// ---------------------------------------------------------------------------------------------
// sin / cos: num-traits "libm" feature (Cargo.toml:17-20) -> Rust `libm` crate, a port of
// musl/FreeBSD msun (Sun fdlibm): k_sin.c, k_cos.c, s_sin.c, s_cos.c, e_rem_pio2.c.  Restated
// from the published algorithm.  |x| >= 2^20*pi/2 (Payne-Hanek in fdlibm) is folded with an
// exact fmod by the double nearest 2*pi instead — a diverged solve, documented in DESIGN.md.
// ---------------------------------------------------------------------------------------------
// @region sincos
ICP_DEV double k_sin(double x, double y, int iy)
{
    const double S1 = bits_to_f64(ICP_HEXD(0xBFC55555u, 0x55555549u));
    const double S2 = bits_to_f64(ICP_HEXD(0x3F811111u, 0x1110F8A6u));
    const double S3 = bits_to_f64(ICP_HEXD(0xBF2A01A0u, 0x19C161D5u));
    const double S4 = bits_to_f64(ICP_HEXD(0x3EC71DE3u, 0x57B1FE7Du));
    const double S5 = bits_to_f64(ICP_HEXD(0xBE5AE5E6u, 0x8A2B9CEBu));
    const double S6 = bits_to_f64(ICP_HEXD(0x3DE5D93Au, 0x5ACFD57Cu));
    double z = x * x;
    double w = z * z;
    double r = S2 + z * (S3 + z * S4) + z * w * (S5 + z * S6);
    double v = z * x;
    if (iy == 0) return x + v * (S1 + z * r);
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

ICP_DEV double k_cos(double x, double y)
{
    const double C1 = bits_to_f64(ICP_HEXD(0x3FA55555u, 0x5555554Cu));
    const double C2 = bits_to_f64(ICP_HEXD(0xBF56C16Cu, 0x16C15177u));
    const double C3 = bits_to_f64(ICP_HEXD(0x3EFA01A0u, 0x19CB1590u));
    const double C4 = bits_to_f64(ICP_HEXD(0xBE927E4Fu, 0x809C52ADu));
    const double C5 = bits_to_f64(ICP_HEXD(0x3E21EE9Eu, 0xBDB4B1C4u));
    const double C6 = bits_to_f64(ICP_HEXD(0xBDA8FAE9u, 0xBE8838D4u));
    double z = x * x;
    double w = z * z;
    double r = z * (C1 + z * (C2 + z * C3)) + (w * w) * (C4 + z * (C5 + z * C6));
    double hz = 0.5 * z;
    w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}

// argument reduction for pi/4 < |x| < 2^20*pi/2 (cold path: ICP updates are milliradians)
ICP_DEV_NOINLINE int rem_pio2(double x, double* y0, double* y1)
{
    const double toint = 6755399441055744.0;
    const double invpio2 = bits_to_f64(ICP_HEXD(0x3FE45F30u, 0x6DC9C883u));
    const double pio2_1 = bits_to_f64(ICP_HEXD(0x3FF921FBu, 0x54400000u));
    const double pio2_1t = bits_to_f64(ICP_HEXD(0x3DD0B461u, 0x1A626331u));
    const double pio2_2 = bits_to_f64(ICP_HEXD(0x3DD0B461u, 0x1A600000u));
    const double pio2_2t = bits_to_f64(ICP_HEXD(0x3BA3198Au, 0x2E037073u));
    const double pio2_3 = bits_to_f64(ICP_HEXD(0x3BA3198Au, 0x2E000000u));
    const double pio2_3t = bits_to_f64(ICP_HEXD(0x397B839Au, 0x252049C1u));
    uint64_t ui = f64_to_bits(x);
    bool neg = (ui >> 63) != 0;
    uint32_t ix = (uint32_t)(ui >> 32) & 0x7fffffffu;
    bool medium = false;
    int k = 0;
    if (ix <= 0x400f6a7au) {  // |x| ~<= 5pi/4
        if ((ix & 0xfffffu) == 0x921fbu) medium = true;
        else k = (ix <= 0x4002d97cu) ? 1 : 2;
    } else if (ix <= 0x401c463bu) {  // |x| ~<= 9pi/4
        if (ix <= 0x4015fdbcu) {
            if (ix == 0x4012d97cu) medium = true;
            else k = 3;
        } else {
            if (ix == 0x401921fbu) medium = true;
            else k = 4;
        }
    } else {
        medium = true;
    }
    if (!medium) {
        double kd = (double)k;  // k*pio2_1 and k*pio2_1t are exact products for k = 1..4
        double a = kd * pio2_1, at = kd * pio2_1t;
        if (!neg) {
            double z = x - a;
            double r0 = z - at;
            *y0 = r0;
            *y1 = (z - r0) - at;
            return k;
        }
        double z = x + a;
        double r0 = z + at;
        *y0 = r0;
        *y1 = (z - r0) + at;
        return -k;
    }
    double fn = x * invpio2 + toint - toint;
    int n = (int)fn;
    double r = x - fn * pio2_1;
    double w = fn * pio2_1t;
    double r0 = r - w;
    int ey = (int)((f64_to_bits(r0) >> 52) & 0x7ff);
    int ex = (int)(ix >> 20);
    if (ex - ey > 16) {
        double t = r;
        w = fn * pio2_2;
        r = t - w;
        w = fn * pio2_2t - ((t - r) - w);
        r0 = r - w;
        ey = (int)((f64_to_bits(r0) >> 52) & 0x7ff);
        if (ex - ey > 49) {
            t = r;
            w = fn * pio2_3;
            r = t - w;
            w = fn * pio2_3t - ((t - r) - w);
            r0 = r - w;
        }
    }
    *y0 = r0;
    *y1 = (r - r0) - w;
    return n;
}

ICP_DEV_NOINLINE void sincos_slow(double x, double* s, double* c)
{
    uint32_t ix = (uint32_t)(f64_to_bits(x) >> 32) & 0x7fffffffu;
    if (ix >= 0x7ff00000u) {  // Inf / NaN
        *s = x - x;
        *c = x - x;
        return;
    }
    if (ix >= 0x413921fbu) {
        const double two_pi = bits_to_f64(ICP_HEXD(0x401921FBu, 0x54442D18u));
        x = fmod(x, two_pi);
        ix = (uint32_t)(f64_to_bits(x) >> 32) & 0x7fffffffu;
        if (ix <= 0x3fe921fbu) {
            *s = ix < 0x3e500000u ? x : k_sin(x, 0.0, 0);
            *c = ix < 0x3e46a09eu ? 1.0 : k_cos(x, 0.0);
            return;
        }
    }
    double y0, y1;
    int n = rem_pio2(x, &y0, &y1);
    double ks = k_sin(y0, y1, 1), kc = k_cos(y0, y1);
    switch (n & 3) {
        case 0: *s = ks; *c = kc; break;
        case 1: *s = kc; *c = -ks; break;
        case 2: *s = -ks; *c = -kc; break;
        default: *s = -kc; *c = ks; break;
    }
}

// f64::sin / f64::cos of the same argument (src/so2.rs:13-14, src/se2.rs:26-27)
ICP_DEV void sincos_libm(double x, double* s, double* c)
{
    uint32_t ix = (uint32_t)(f64_to_bits(x) >> 32) & 0x7fffffffu;
    if (ix <= 0x3fe921fbu) {  // |x| ~< pi/4
        *s = ix < 0x3e500000u ? x : k_sin(x, 0.0, 0);
        *c = ix < 0x3e46a09eu ? 1.0 : k_cos(x, 0.0);
        return;
    }
    sincos_slow(x, s, c);
}

// ---------------------------------------------------------------------------------------------
// src/huber.rs:6-15 / 17-26 (argument e is the SQUARED residual)
// @region huber
// Branch-free on purpose: lanes of a warp rarely agree on the branch, and without it the square
// roots / quotients of a thread's points are independent instructions the scheduler interleaves.
ICP_DEV double huber_rho(double e)
{
    const double k2 = kHuberK * kHuberK;
    const double far = 2. * kHuberK * sqrt(e) - k2;
    return (e <= k2) ? e : far;
}
ICP_DEV double huber_drho(double e)
{
    const double k2 = kHuberK * kHuberK;
    if (e <= k2) return 1.;
    return kHuberK / sqrt(e);
}

// huber_drho(r * r) for a residual component r (lib.rs:250).  In radix 2 with round-to-nearest,
// sqrt(fl(r*r)) == |r| exactly whenever r*r neither underflows nor overflows (Boldo, "Stupid is
// as stupid does: taking the square root of the square of a floating-point number", 2015), so
// the square root is skipped: k / sqrt(e) == k / |r|.  The branch e <= k*k already covers the
// underflow side; an overflowing e keeps the literal formula.
ICP_DEV double huber_drho_of_residual(double r)
{
    const double k2 = kHuberK * kHuberK;
    const double e = r * r;
    const double den = (e > kF64Max) ? e : fabs(r);   // e == +inf there and sqrt(+inf) == +inf
    const double q = kHuberK / den;                   // r == 0: k / 0 = +inf, not selected
    return (e <= k2) ? 1. : q;
}

// ---------------------------------------------------------------------------------------------
// src/transform.rs:22-24: rot * p + t (nalgebra gemv sums columns in ascending order)
// @region transform_ops
ICP_DEV void transform_apply(const Transform& T, double x, double y, double* ox, double* oy)
{
    *ox = (T.r00 * x + T.r01 * y) + T.tx;
    *oy = (T.r10 * x + T.r11 * y) + T.ty;
}

// src/transform.rs:42-51: A * B
ICP_DEV Transform transform_mul(const Transform& A, const Transform& B)
{
    Transform o;
    o.r00 = A.r00 * B.r00 + A.r01 * B.r10;
    o.r01 = A.r00 * B.r01 + A.r01 * B.r11;
    o.r10 = A.r10 * B.r00 + A.r11 * B.r10;
    o.r11 = A.r10 * B.r01 + A.r11 * B.r11;
    o.tx = (A.r00 * B.tx + A.r01 * B.ty) + A.tx;
    o.ty = (A.r10 * B.tx + A.r11 * B.ty) + A.ty;
    return o;
}

ICP_DEV Transform transform_identity()  // src/transform.rs:34-39
{
    Transform o;
    o.r00 = 1.; o.r01 = 0.; o.r10 = 0.; o.r11 = 1.; o.tx = 0.; o.ty = 0.;
    return o;
}

// Transform::new(param) -> se2::calc_rt -> so2::new_rotation2
// (src/transform.rs:13-16, src/se2.rs:21-41, src/so2.rs:8-17)
// @region transform_new
ICP_DEV Transform transform_new(double vx, double vy, double theta)
{
    double s, c;
    sincos_libm(theta, &s, &c);
    Transform o;
    o.r00 = c; o.r01 = -s; o.r10 = s; o.r11 = c;
    if (theta == 0.) {
        o.tx = vx;
        o.ty = vy;
    } else {
        o.tx = (s * vx - (1. - c) * vy) / theta;
        o.ty = ((1. - c) * vx + s * vy) / theta;
    }
    return o;
}

// src/linalg.rs:3-29 followed by `-jtj_inv * jtr` (src/lib.rs:258).  m is row-major jtj.
// Returns false for None (det == 0 exactly).
// @region solve3x3
ICP_DEV bool solve_normal_equations(const double m[9], const double jtr[3], double delta[3])
{
    double m00 = m[0], m01 = m[1], m02 = m[2];
    double m10 = m[3], m11 = m[4], m12 = m[5];
    double m20 = m[6], m21 = m[7], m22 = m[8];
    double det = m00 * (m22 * m11 - m21 * m12) - m10 * (m22 * m01 - m21 * m02) +
                 m20 * (m12 * m01 - m11 * m02);
    if (det == 0.) return false;
    double i00 = (m22 * m11 - m21 * m12) / det;
    double i01 = (-(m22 * m01 - m21 * m02)) / det;
    double i02 = (m12 * m01 - m11 * m02) / det;
    double i10 = (-(m22 * m10 - m20 * m12)) / det;
    double i11 = (m22 * m00 - m20 * m02) / det;
    double i12 = (-(m12 * m00 - m10 * m02)) / det;
    double i20 = (m21 * m10 - m20 * m11) / det;
    double i21 = (-(m21 * m00 - m20 * m01)) / det;
    double i22 = (m11 * m00 - m10 * m01) / det;
    delta[0] = ((-i00) * jtr[0] + (-i01) * jtr[1]) + (-i02) * jtr[2];
    delta[1] = ((-i10) * jtr[0] + (-i11) * jtr[1]) + (-i12) * jtr[2];
    delta[2] = ((-i20) * jtr[0] + (-i21) * jtr[1]) + (-i22) * jtr[2];
    return true;
}

// Order-preserving map f64 -> u64 (total order; -0.0 sorts just below +0.0, which cannot change
// a median VALUE used downstream: see DESIGN.md "median").
// @region misc_math
ICP_DEV uint64_t f64_order_key(double v)
{
    uint64_t b = f64_to_bits(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
ICP_DEV double f64_from_order_key(uint64_t k)
{
    return bits_to_f64((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k);
}

// correspondence checksum: splitmix64 finaliser (definition in DESIGN.md "correspondence checksum")
ICP_DEV uint64_t mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

}  // namespace icp
