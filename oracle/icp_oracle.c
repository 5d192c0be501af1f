/*
 * icp_oracle.c — CPU ORACLE (test infrastructure, see icp_oracle.h for the rules).
 *
 * Plain-C restatement of tier4/icp_rust's ICP iteration loop.  Every function cites the
 * reference file:line (relative to /root/reference) it follows.  Build with
 *   gcc -O2 -ffp-contract=off -fno-fast-math   (see oracle/Makefile)
 * so that every expression is evaluated as separately rounded IEEE-754 binary64 operations
 * in source order, as rustc does.
 */
#include "icp_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* =====================================================================================
 * sin / cos.  The crate is no_std by default (Cargo.toml:6-8, src/lib.rs:3) and pulls
 * f64::sin / f64::cos from num_traits::real::Real with feature "libm" (Cargo.toml:17-20;
 * used at src/so2.rs:13-14, src/se2.rs:26-27), i.e. the Rust `libm` crate, a port of
 * musl libc's math library (FreeBSD msun / Sun fdlibm).  That source is not under
 * /root/reference (crates.io dependency, version unpinned: no Cargo.lock), so the
 * published fdlibm algorithm is restated here: k_sin.c, k_cos.c, s_sin.c, s_cos.c and the
 * non-huge branches of e_rem_pio2.c.  For |x| >= 2^20*pi/2 fdlibm switches to Payne-Hanek
 * (__rem_pio2_large); an SE(2) update that large is a diverged solve, so this oracle
 * (and the CUDA path) instead folds the argument with an exact fmod by the double nearest
 * 2*pi first.  Documented deviation, irrelevant to any physical input.
 * ===================================================================================== */
static double from_bits(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static uint64_t to_bits(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }

#define HEXD(hi, lo) (((uint64_t)(hi) << 32) | (uint64_t)(lo))

static double k_sin(double x, double y, int iy)
{
    const double S1 = from_bits(HEXD(0xBFC55555u, 0x55555549u)); /* -1.66666666666666324348e-01 */
    const double S2 = from_bits(HEXD(0x3F811111u, 0x1110F8A6u)); /*  8.33333333332248946124e-03 */
    const double S3 = from_bits(HEXD(0xBF2A01A0u, 0x19C161D5u)); /* -1.98412698298579493134e-04 */
    const double S4 = from_bits(HEXD(0x3EC71DE3u, 0x57B1FE7Du)); /*  2.75573137070700676789e-06 */
    const double S5 = from_bits(HEXD(0xBE5AE5E6u, 0x8A2B9CEBu)); /* -2.50507602534068634195e-08 */
    const double S6 = from_bits(HEXD(0x3DE5D93Au, 0x5ACFD57Cu)); /*  1.58969099521155010221e-10 */
    double z = x * x;
    double w = z * z;
    double r = S2 + z * (S3 + z * S4) + z * w * (S5 + z * S6);
    double v = z * x;
    if (iy == 0)
        return x + v * (S1 + z * r);
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

static double k_cos(double x, double y)
{
    const double C1 = from_bits(HEXD(0x3FA55555u, 0x5555554Cu)); /*  4.16666666666666019037e-02 */
    const double C2 = from_bits(HEXD(0xBF56C16Cu, 0x16C15177u)); /* -1.38888888888741095749e-03 */
    const double C3 = from_bits(HEXD(0x3EFA01A0u, 0x19CB1590u)); /*  2.48015872894767294178e-05 */
    const double C4 = from_bits(HEXD(0xBE927E4Fu, 0x809C52ADu)); /* -2.75573143513906633035e-07 */
    const double C5 = from_bits(HEXD(0x3E21EE9Eu, 0xBDB4B1C4u)); /*  2.08757232129817482790e-09 */
    const double C6 = from_bits(HEXD(0xBDA8FAE9u, 0xBE8838D4u)); /* -1.13596475577881948265e-11 */
    double z = x * x;
    double w = z * z;
    double r = z * (C1 + z * (C2 + z * C3)) + (w * w) * (C4 + z * (C5 + z * C6));
    double hz = 0.5 * z;
    w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + (z * r - x * y));
}

/* e_rem_pio2.c without the Payne-Hanek branch; caller guarantees |x| < 2^20*pi/2. */
static int rem_pio2(double x, double *y)
{
    const double toint = 6755399441055744.0;                          /* 1.5 / DBL_EPSILON */
    const double invpio2 = from_bits(HEXD(0x3FE45F30u, 0x6DC9C883u)); /* 6.36619772367581382433e-01 */
    const double pio2_1 = from_bits(HEXD(0x3FF921FBu, 0x54400000u));  /* 1.57079632673412561417e+00 */
    const double pio2_1t = from_bits(HEXD(0x3DD0B461u, 0x1A626331u)); /* 6.07710050650619224932e-11 */
    const double pio2_2 = from_bits(HEXD(0x3DD0B461u, 0x1A600000u));  /* 6.07710050630396597660e-11 */
    const double pio2_2t = from_bits(HEXD(0x3BA3198Au, 0x2E037073u)); /* 2.02226624879595063154e-21 */
    const double pio2_3 = from_bits(HEXD(0x3BA3198Au, 0x2E000000u));  /* 2.02226624871116645580e-21 */
    const double pio2_3t = from_bits(HEXD(0x397B839Au, 0x252049C1u)); /* 8.47842766036889956997e-32 */
    uint64_t ui = to_bits(x);
    int sign = (int)(ui >> 63);
    uint32_t ix = (uint32_t)(ui >> 32) & 0x7fffffffu;
    double z, w, t, r, fn;
    int n, ex, ey;

    if (ix <= 0x400f6a7au) { /* |x| ~<= 5pi/4 */
        if ((ix & 0xfffffu) == 0x921fbu) /* |x| ~= pi/2 or pi: cancellation, medium path */
            goto medium;
        if (ix <= 0x4002d97cu) { /* |x| ~<= 3pi/4 */
            if (!sign) {
                z = x - pio2_1;
                y[0] = z - pio2_1t;
                y[1] = (z - y[0]) - pio2_1t;
                return 1;
            }
            z = x + pio2_1;
            y[0] = z + pio2_1t;
            y[1] = (z - y[0]) + pio2_1t;
            return -1;
        }
        if (!sign) {
            z = x - 2 * pio2_1;
            y[0] = z - 2 * pio2_1t;
            y[1] = (z - y[0]) - 2 * pio2_1t;
            return 2;
        }
        z = x + 2 * pio2_1;
        y[0] = z + 2 * pio2_1t;
        y[1] = (z - y[0]) + 2 * pio2_1t;
        return -2;
    }
    if (ix <= 0x401c463bu) {     /* |x| ~<= 9pi/4 */
        if (ix <= 0x4015fdbcu) { /* |x| ~<= 7pi/4 */
            if (ix == 0x4012d97cu) /* |x| ~= 3pi/2 */
                goto medium;
            if (!sign) {
                z = x - 3 * pio2_1;
                y[0] = z - 3 * pio2_1t;
                y[1] = (z - y[0]) - 3 * pio2_1t;
                return 3;
            }
            z = x + 3 * pio2_1;
            y[0] = z + 3 * pio2_1t;
            y[1] = (z - y[0]) + 3 * pio2_1t;
            return -3;
        }
        if (ix == 0x401921fbu) /* |x| ~= 2pi */
            goto medium;
        if (!sign) {
            z = x - 4 * pio2_1;
            y[0] = z - 4 * pio2_1t;
            y[1] = (z - y[0]) - 4 * pio2_1t;
            return 4;
        }
        z = x + 4 * pio2_1;
        y[0] = z + 4 * pio2_1t;
        y[1] = (z - y[0]) + 4 * pio2_1t;
        return -4;
    }
medium:
    fn = x * invpio2 + toint - toint; /* rint(x/(pi/2)) */
    n = (int)fn;
    r = x - fn * pio2_1;
    w = fn * pio2_1t; /* 1st round, good to 85 bits */
    y[0] = r - w;
    ey = (int)((to_bits(y[0]) >> 52) & 0x7ff);
    ex = (int)(ix >> 20);
    if (ex - ey > 16) { /* 2nd round, good to 118 bits */
        t = r;
        w = fn * pio2_2;
        r = t - w;
        w = fn * pio2_2t - ((t - r) - w);
        y[0] = r - w;
        ey = (int)((to_bits(y[0]) >> 52) & 0x7ff);
        if (ex - ey > 49) { /* 3rd round, good to 151 bits */
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

static double fold_huge(double x)
{
    /* documented deviation from fdlibm: exact fmod by the double nearest 2*pi */
    const double two_pi = from_bits(HEXD(0x401921FBu, 0x54442D18u));
    return fmod(x, two_pi);
}

double orc_sin(double x)
{
    uint32_t ix = (uint32_t)(to_bits(x) >> 32) & 0x7fffffffu;
    double y[2];
    int n;
    if (ix <= 0x3fe921fbu) {   /* |x| ~< pi/4 */
        if (ix < 0x3e500000u)  /* |x| < 2**-26 */
            return x;
        return k_sin(x, 0.0, 0);
    }
    if (ix >= 0x7ff00000u) /* Inf or NaN */
        return x - x;
    if (ix >= 0x413921fbu) {
        x = fold_huge(x);
        ix = (uint32_t)(to_bits(x) >> 32) & 0x7fffffffu;
        if (ix <= 0x3fe921fbu)
            return ix < 0x3e500000u ? x : k_sin(x, 0.0, 0);
    }
    n = rem_pio2(x, y);
    switch (n & 3) {
    case 0: return k_sin(y[0], y[1], 1);
    case 1: return k_cos(y[0], y[1]);
    case 2: return -k_sin(y[0], y[1], 1);
    default: return -k_cos(y[0], y[1]);
    }
}

double orc_cos(double x)
{
    uint32_t ix = (uint32_t)(to_bits(x) >> 32) & 0x7fffffffu;
    double y[2];
    int n;
    if (ix <= 0x3fe921fbu) {   /* |x| ~< pi/4 */
        if (ix < 0x3e46a09eu)  /* |x| < 2**-27 * sqrt(2) */
            return 1.0;
        return k_cos(x, 0.0);
    }
    if (ix >= 0x7ff00000u)
        return x - x;
    if (ix >= 0x413921fbu) {
        x = fold_huge(x);
        ix = (uint32_t)(to_bits(x) >> 32) & 0x7fffffffu;
        if (ix <= 0x3fe921fbu)
            return ix < 0x3e46a09eu ? 1.0 : k_cos(x, 0.0);
    }
    n = rem_pio2(x, y);
    switch (n & 3) {
    case 0: return k_cos(y[0], y[1]);
    case 1: return -k_sin(y[0], y[1], 1);
    case 2: return -k_cos(y[0], y[1]);
    default: return k_sin(y[0], y[1], 1);
    }
}

/* ===================================================================================== */
/* src/huber.rs:6-15 — rho on the SQUARED argument e */
double orc_huber_rho(double e, double k)
{
    double k_squared = k * k;
    if (e <= k_squared)
        return e;
    return 2. * k * sqrt(e) - k_squared;
}

/* src/huber.rs:17-26 */
double orc_huber_drho(double e, double k)
{
    double k_squared = k * k;
    if (e <= k_squared)
        return 1.;
    return k / sqrt(e);
}

/* =====================================================================================
 * src/stats.rs:11-28 mutable_median.
 * The reference calls select_nth_unstable_by(n/2-1) then select_nth_unstable_by(n/2) on
 * the whole vector and reads both slots.  Restated here as the two order statistics those
 * calls are written to produce (and that test_mutable_median, stats.rs:69-90, asserts).
 * NOTE (unpinned, see DESIGN.md "median"): the second select may legally permute the lower
 * half, so with rustc's pdqselect `input[n/2-1]` can be a non-maximal lower-half element
 * for large even n; that behaviour depends on the (unpinned nightly) toolchain's
 * core::slice::select internals and is not reproduced.
 * ===================================================================================== */
static void swap_d(double *a, double *b) { double t = *a; *a = *b; *b = t; }

/* places the k-th smallest (0-based) at buf[k]; everything before is <=, after is >= */
static void nth_element(double *buf, size_t n, size_t k)
{
    ptrdiff_t lo = 0, hi = (ptrdiff_t)n - 1, kk = (ptrdiff_t)k;
    while (lo < hi) {
        ptrdiff_t mid = lo + (hi - lo) / 2, i = lo, j = hi;
        double pivot;
        if (buf[mid] < buf[lo]) swap_d(&buf[mid], &buf[lo]);
        if (buf[hi] < buf[lo]) swap_d(&buf[hi], &buf[lo]);
        if (buf[hi] < buf[mid]) swap_d(&buf[hi], &buf[mid]);
        pivot = buf[mid];
        while (i <= j) {
            while (buf[i] < pivot) i++;
            while (pivot < buf[j]) j--;
            if (i <= j) {
                swap_d(&buf[i], &buf[j]);
                i++;
                j--;
            }
        }
        /* now [lo..j] <= pivot <= [i..hi], j < i */
        if (kk <= j) hi = j;
        else if (kk >= i) lo = i;
        else return;
    }
}

int orc_median(double *buf, size_t n, double *out)
{
    if (n == 0)
        return 0; /* stats.rs:15-17 */
    if (n % 2 == 1) {
        nth_element(buf, n, n / 2); /* stats.rs:18-21 */
        *out = buf[n / 2];
        return 1;
    }
    nth_element(buf, n, n / 2 - 1); /* stats.rs:23 */
    {
        /* stats.rs:24: the (n/2)-th order statistic = min of the upper part */
        size_t i, arg = n / 2;
        for (i = n / 2 + 1; i < n; i++)
            if (buf[i] < buf[arg]) arg = i;
        swap_d(&buf[n / 2], &buf[arg]);
    }
    {
        double b = buf[n / 2 - 1]; /* stats.rs:25 */
        double c = buf[n / 2];     /* stats.rs:26 */
        *out = (b + c) / 2.;       /* stats.rs:27 */
    }
    return 1;
}

/* src/stats.rs:30-37 */
int orc_mad(double *buf, size_t n, double *out)
{
    double m;
    size_t i;
    int ok;
    double *a;
    if (!orc_median(buf, n, &m))
        return 0;
    a = (double *)malloc(n * sizeof(double));
    for (i = 0; i < n; i++)
        a[i] = fabs(buf[i] - m);
    ok = orc_median(a, n, out);
    free(a);
    return ok;
}

/* src/stats.rs:39-47 */
int orc_standard_deviation(double *buf, size_t n, double *out)
{
    const double ppf34 = 1.482602218505602; /* stats.rs:42 */
    double m;
    if (!orc_mad(buf, n, &m))
        return 0;
    *out = ppf34 * m;
    return 1;
}

/* src/stats.rs:49-60 */
int orc_calc_stddevs(const double *residuals, size_t n, int dim, double *out)
{
    int j;
    size_t i;
    double *col = (double *)malloc((n ? n : 1) * sizeof(double));
    for (j = 0; j < dim; j++) {
        double s;
        for (i = 0; i < n; i++)
            col[i] = residuals[i * (size_t)dim + (size_t)j];
        if (!orc_standard_deviation(col, n, &s)) {
            free(col);
            return 0;
        }
        out[j] = s;
    }
    free(col);
    return 1;
}

/* ===================================================================================== */
/* src/linalg.rs:3-29 */
int orc_inverse3x3(const double m[9], double out[9])
{
    double m00 = m[0], m01 = m[1], m02 = m[2];
    double m10 = m[3], m11 = m[4], m12 = m[5];
    double m20 = m[6], m21 = m[7], m22 = m[8];
    double det = m00 * (m22 * m11 - m21 * m12) - m10 * (m22 * m01 - m21 * m02) +
                 m20 * (m12 * m01 - m11 * m02); /* linalg.rs:15-17 */
    if (det == 0.)
        return 0; /* linalg.rs:18-20 */
    out[0] = (m22 * m11 - m21 * m12) / det;
    out[1] = (-(m22 * m01 - m21 * m02)) / det;
    out[2] = (m12 * m01 - m11 * m02) / det;
    out[3] = (-(m22 * m10 - m20 * m12)) / det;
    out[4] = (m22 * m00 - m20 * m02) / det;
    out[5] = (-(m12 * m00 - m10 * m02)) / det;
    out[6] = (m21 * m10 - m20 * m11) / det;
    out[7] = (-(m21 * m00 - m20 * m01)) / det;
    out[8] = (m11 * m00 - m10 * m01) / det;
    return 1;
}

/* ===================================================================================== */
/* src/so2.rs:8-17 */
void orc_so2_new_rotation2(double theta, double rot[4])
{
    rot[0] = orc_cos(theta);
    rot[1] = -orc_sin(theta);
    rot[2] = orc_sin(theta);
    rot[3] = orc_cos(theta);
}

/* src/so2.rs:19-21 (off the hot path; system atan2) */
double orc_so2_log(const double rot[4]) { return atan2(rot[2], rot[0]); }

/* src/so2.rs:23-31 */
void orc_so2_exp(double theta, double rot[4])
{
    double c = orc_cos(theta), s = orc_sin(theta);
    rot[0] = c;
    rot[1] = -s;
    rot[2] = s;
    rot[3] = c;
}

/* src/se2.rs:11-19 */
void orc_se2_get_rt(const double m[9], double rot[4], double t[2])
{
    rot[0] = m[0]; rot[1] = m[1];
    rot[2] = m[3]; rot[3] = m[4];
    t[0] = m[2];
    t[1] = m[5];
}

/* src/se2.rs:21-41 */
void orc_se2_calc_rt(const double param[3], double rot[4], double t[2])
{
    double theta = param[2];
    double c, s, vx, vy;
    orc_so2_new_rotation2(theta, rot);
    c = orc_cos(theta);
    s = orc_sin(theta);
    vx = param[0];
    vy = param[1];
    if (theta == 0.) {
        t[0] = vx;
        t[1] = vy;
    } else {
        t[0] = (s * vx - (1. - c) * vy) / theta;
        t[1] = ((1. - c) * vx + s * vy) / theta;
    }
}

/* src/se2.rs:43-52 */
void orc_se2_exp(const double param[3], double out[9])
{
    double rot[4], t[2];
    orc_se2_calc_rt(param, rot, t);
    out[0] = rot[0]; out[1] = rot[1]; out[2] = t[0];
    out[3] = rot[2]; out[4] = rot[3]; out[5] = t[1];
    out[6] = 0.; out[7] = 0.; out[8] = 1.;
}

/* src/se2.rs:54-77 (off the hot path) */
void orc_se2_log(const double m[9], double out[3])
{
    double rot[4], t[2], v[4];
    double theta;
    orc_se2_get_rt(m, rot, t);
    theta = orc_so2_log(rot);
    if (theta == 0.) {
        v[0] = 1.; v[1] = 0.; v[2] = 0.; v[3] = 1.;
    } else if (theta == 3.14159265358979323846) {
        v[0] = 0.; v[1] = 0.5 * theta; v[2] = -0.5 * theta; v[3] = 0.;
    } else {
        double k = orc_sin(theta) / (1. - orc_cos(theta));
        v[0] = (0.5 * theta) * k; v[1] = (0.5 * theta) * 1.;
        v[2] = (0.5 * theta) * -1.; v[3] = (0.5 * theta) * k;
    }
    out[0] = v[0] * t[0] + v[1] * t[1];
    out[1] = v[2] * t[0] + v[3] * t[1];
    out[2] = theta;
}

/* ===================================================================================== */
/* src/transform.rs:13-16 */
void orc_transform_new(const double param[3], double T[6]) { orc_se2_calc_rt(param, T, T + 4); }

/* src/transform.rs:34-39 */
void orc_transform_identity(double T[6])
{
    T[0] = 1.; T[1] = 0.; T[2] = 0.; T[3] = 1.; T[4] = 0.; T[5] = 0.;
}

/* src/transform.rs:22-24: rot * landmark + t (nalgebra gemv: column-ascending sum) */
void orc_transform_apply(const double T[6], const double p[2], double out[2])
{
    double x = p[0], y = p[1];
    out[0] = (T[0] * x + T[1] * y) + T[4];
    out[1] = (T[2] * x + T[3] * y) + T[5];
}

/* src/transform.rs:26-32: inverse rotation = transpose; t = -(inv_rot * t) */
void orc_transform_inverse(const double T[6], double out[6])
{
    double i00 = T[0], i01 = T[2], i10 = T[1], i11 = T[3];
    out[0] = i00; out[1] = i01; out[2] = i10; out[3] = i11;
    out[4] = -(i00 * T[4] + i01 * T[5]);
    out[5] = -(i10 * T[4] + i11 * T[5]);
}

/* src/transform.rs:42-51: rot = A.rot*B.rot (no re-orthonormalisation), t = A.rot*B.t + A.t */
void orc_transform_mul(const double A[6], const double B[6], double out[6])
{
    double r00 = A[0] * B[0] + A[1] * B[2];
    double r01 = A[0] * B[1] + A[1] * B[3];
    double r10 = A[2] * B[0] + A[3] * B[2];
    double r11 = A[2] * B[1] + A[3] * B[3];
    double tx = (A[0] * B[4] + A[1] * B[5]) + A[4];
    double ty = (A[2] * B[4] + A[3] * B[5]) + A[5];
    out[0] = r00; out[1] = r01; out[2] = r10; out[3] = r11; out[4] = tx; out[5] = ty;
}

/* src/norm.rs:8-21 */
double orc_norm(const double *v, size_t n)
{
    double res = 0.;
    size_t i;
    for (i = 0; i < n; i++)
        res += v[i] * v[i];
    return sqrt(res);
}

/* ===================================================================================== */
/* src/lib.rs:34-36 */
void orc_residual(const double T[6], const double s[2], const double d[2], double out[2])
{
    double p[2];
    orc_transform_apply(T, s, p);
    out[0] = p[0] - d[0];
    out[1] = p[1] - d[1];
}

/* src/lib.rs:38-43 */
double orc_error(const double T[6], const double *src, const double *dst, size_t n)
{
    double sum = 0.;
    size_t i;
    for (i = 0; i < n; i++) {
        double r[2];
        orc_residual(T, src + 2 * i, dst + 2 * i, r);
        sum = sum + (r[0] * r[0] + r[1] * r[1]);
    }
    return sum;
}

/* src/lib.rs:45-50 */
double orc_huber_error(const double T[6], const double *src, const double *dst, size_t n)
{
    double sum = 0.;
    size_t i;
    for (i = 0; i < n; i++) {
        double r[2];
        orc_residual(T, src + 2 * i, dst + 2 * i, r);
        sum = sum + orc_huber_rho(r[0] * r[0] + r[1] * r[1], ORC_HUBER_K);
    }
    return sum;
}

/* src/lib.rs:176-184: J = [[r00 r01 b0],[r10 r11 b1]], b = rot * (-ly, lx) */
static void jacobian(const double T[6], const double l[2], double J[6])
{
    double a0 = -l[1], a1 = l[0];
    double b0 = T[0] * a0 + T[1] * a1;
    double b1 = T[2] * a0 + T[3] * a1;
    J[0] = T[0]; J[1] = T[1]; J[2] = b0;
    J[3] = T[2]; J[4] = T[3]; J[5] = b1;
}

/* src/lib.rs:186-189: len > 0 && len >= 2 */
static int check_input_size(size_t n) { return n > 0 && n >= 2; }

/* -inv * jtr with nalgebra's column-ascending gemv (lib.rs:213,258) */
static void solve_neg(const double inv[9], const double jtr[3], double delta[3])
{
    int i;
    for (i = 0; i < 3; i++)
        delta[i] = ((-inv[3 * i + 0]) * jtr[0] + (-inv[3 * i + 1]) * jtr[1]) +
                   (-inv[3 * i + 2]) * jtr[2];
}

/* src/lib.rs:191-216 (public, unweighted; not on the estimate path, kept for the golden tests) */
int orc_gauss_newton_update(const double T[6], const double *src, const double *dst, size_t n,
                            double delta[3])
{
    double jtr[3] = {0., 0., 0.}, jtj[9] = {0., 0., 0., 0., 0., 0., 0., 0., 0.}, inv[9];
    size_t i;
    int a, b;
    if (!check_input_size(n))
        return 0;
    for (i = 0; i < n; i++) {
        double J[6], r[2];
        jacobian(T, src + 2 * i, J);
        orc_residual(T, src + 2 * i, dst + 2 * i, r);
        for (a = 0; a < 3; a++) {
            /* J^T r : row a of J^T times r, k-ascending */
            double v = J[a] * r[0] + J[3 + a] * r[1];
            jtr[a] = jtr[a] + v;
            for (b = 0; b < 3; b++) {
                double h = J[a] * J[b] + J[3 + a] * J[3 + b];
                jtj[3 * a + b] = jtj[3 * a + b] + h;
            }
        }
    }
    if (!orc_inverse3x3(jtj, inv))
        return 0;
    solve_neg(inv, jtr, delta);
    return 1;
}

/* src/lib.rs:218-261 */
int orc_weighted_gauss_newton_update(const double T[6], const double *src, const double *dst,
                                     size_t n, double delta[3])
{
    double jtr[3] = {0., 0., 0.}, jtj[9] = {0., 0., 0., 0., 0., 0., 0., 0., 0.}, inv[9];
    double stddevs[2];
    double *res;
    size_t i;
    int j, a, b;
    if (!check_input_size(n))
        return 0; /* lib.rs:225-228 */
    res = (double *)malloc(2 * n * sizeof(double));
    for (i = 0; i < n; i++) /* lib.rs:230-234 */
        orc_residual(T, src + 2 * i, dst + 2 * i, res + 2 * i);
    if (!orc_calc_stddevs(res, n, 2, stddevs)) { /* lib.rs:236-238 */
        free(res);
        return 0;
    }
    for (i = 0; i < n; i++) { /* lib.rs:242-255 */
        double J[6];
        jacobian(T, src + 2 * i, J);
        for (j = 0; j < 2; j++) {
            double g, r_ij, w_ij, c;
            const double *Jj = J + 3 * j;
            if (stddevs[j] == 0.)
                continue; /* lib.rs:245-247 */
            g = 1. / stddevs[j];
            r_ij = res[2 * i + (size_t)j];
            w_ij = orc_huber_drho(r_ij * r_ij, ORC_HUBER_K);
            c = w_ij * g;
            for (a = 0; a < 3; a++) {
                double cj = c * Jj[a];
                jtr[a] = jtr[a] + cj * r_ij; /* lib.rs:252 */
                for (b = 0; b < 3; b++)
                    jtj[3 * a + b] = jtj[3 * a + b] + cj * Jj[b]; /* lib.rs:253 */
            }
        }
    }
    free(res);
    if (!orc_inverse3x3(jtj, inv)) /* lib.rs:257-260 */
        return 0;
    solve_neg(inv, jtr, delta);
    return 1;
}

/* src/lib.rs:59-84 */
/* margins (nullable, test diagnostics only): [0] = smallest |delta.delta - 1e-6| / 1e-6 and
 * [1] = smallest |error - prev_error| / max(error, prev_error) over the decisions taken; the
 * caller initialises both to +inf.  They say how close a break test came to its threshold. */
static void estimate_transform_ex(const double *src, const double *dst, size_t n, double T[6],
                                  uint32_t *accepted, uint32_t *reason, double *margins)
{
    const double delta_norm_threshold = 1e-6; /* lib.rs:60 */
    const uint32_t max_iter = 200;            /* lib.rs:61 */
    double prev_error = DBL_MAX;              /* lib.rs:63 */
    uint32_t it, acc = 0, why = ORC_BREAK_CAP;
    orc_transform_identity(T); /* lib.rs:65 */
    for (it = 0; it < max_iter; it++) {
        double delta[3], error, E[6], Tn[6];
        if (!orc_weighted_gauss_newton_update(T, src, dst, n, delta)) {
            why = ORC_BREAK_NONE;
            break; /* lib.rs:67-69 */
        }
        if (margins) {
            double dd = (delta[0] * delta[0] + delta[1] * delta[1]) + delta[2] * delta[2];
            double mg = fabs(dd - delta_norm_threshold) / delta_norm_threshold;
            if (mg < margins[0]) margins[0] = mg;
        }
        if ((delta[0] * delta[0] + delta[1] * delta[1]) + delta[2] * delta[2] <
            delta_norm_threshold) {
            why = ORC_BREAK_DELTA;
            break; /* lib.rs:71-73 */
        }
        error = orc_huber_error(T, src, dst, n); /* lib.rs:75 */
        if (margins && prev_error != DBL_MAX) {
            double big = error > prev_error ? error : prev_error;
            double mg = big > 0. ? fabs(error - prev_error) / big : 0.;
            if (mg < margins[1]) margins[1] = mg;
        }
        if (error > prev_error) {
            why = ORC_BREAK_ERROR;
            break; /* lib.rs:76-78 */
        }
        prev_error = error;
        orc_transform_new(delta, E);
        orc_transform_mul(E, T, Tn); /* lib.rs:81 */
        memcpy(T, Tn, sizeof(Tn));
        acc++;
    }
    if (accepted) *accepted = acc;
    if (reason) *reason = why;
}

void orc_estimate_transform(const double *src, const double *dst, size_t n, double T[6],
                            uint32_t *accepted, uint32_t *reason)
{
    estimate_transform_ex(src, dst, n, T, accepted, reason, NULL);
}

/* =====================================================================================
 * Nearest neighbour.  The reference uses nearest_neighbor::KdTree (external crate, source
 * absent; call sites src/lib.rs:26,99,121,141,164).  Contract adopted (north_star): exact
 * Euclidean 1-NN over d2 = dx*dx + dy*dy (+ dz*dz), lowest index on equal d2.
 * ===================================================================================== */
static double dist2(const double *p, const double *q, int dim)
{
    double dx = q[0] - p[0], dy = q[1] - p[1];
    double d2 = dx * dx + dy * dy;
    if (dim == 3) {
        double dz = q[2] - p[2];
        d2 = d2 + dz * dz;
    }
    return d2;
}

int64_t orc_nn_brute(const double *dst, size_t m, int dim, const double *q)
{
    int64_t best = -1;
    double bd = 0.;
    size_t i;
    for (i = 0; i < m; i++) {
        double d2 = dist2(dst + i * (size_t)dim, q, dim);
        if (best < 0 || d2 < bd) {
            best = (int64_t)i;
            bd = d2;
        }
    }
    return best;
}

struct orc_kdtree {
    const double *pts;
    size_t m;
    int dim;
    int32_t *perm;  /* implicit balanced tree over perm[lo,hi): node = perm[mid] */
    uint8_t *axis;  /* split axis of node stored at position mid */
};

typedef struct { const double *pts; int dim; int axis; } kd_cmp_ctx;

static int kd_less(const kd_cmp_ctx *c, int32_t a, int32_t b)
{
    double va = c->pts[(size_t)a * (size_t)c->dim + (size_t)c->axis];
    double vb = c->pts[(size_t)b * (size_t)c->dim + (size_t)c->axis];
    if (va < vb) return 1;
    if (vb < va) return 0;
    return a < b;
}

static void kd_select(int32_t *p, size_t n, size_t k, const kd_cmp_ctx *c)
{
    ptrdiff_t lo = 0, hi = (ptrdiff_t)n - 1, kk = (ptrdiff_t)k;
    while (lo < hi) {
        ptrdiff_t mid = lo + (hi - lo) / 2, i = lo, j = hi;
        int32_t pivot, t;
        if (kd_less(c, p[mid], p[lo])) { t = p[mid]; p[mid] = p[lo]; p[lo] = t; }
        if (kd_less(c, p[hi], p[lo])) { t = p[hi]; p[hi] = p[lo]; p[lo] = t; }
        if (kd_less(c, p[hi], p[mid])) { t = p[hi]; p[hi] = p[mid]; p[mid] = t; }
        pivot = p[mid];
        while (i <= j) {
            while (kd_less(c, p[i], pivot)) i++;
            while (kd_less(c, pivot, p[j])) j--;
            if (i <= j) {
                t = p[i]; p[i] = p[j]; p[j] = t;
                i++;
                j--;
            }
        }
        if (kk <= j) hi = j;
        else if (kk >= i) lo = i;
        else return;
    }
}

static void kd_build(orc_kdtree *t, size_t lo, size_t hi)
{
    size_t n = hi - lo, mid, i;
    int d, best_axis = 0;
    double best_ext = -1.;
    kd_cmp_ctx c;
    if (n == 0) return;
    mid = lo + n / 2;
    for (d = 0; d < t->dim; d++) { /* widest-extent axis */
        double mn = DBL_MAX, mx = -DBL_MAX;
        for (i = lo; i < hi; i++) {
            double v = t->pts[(size_t)t->perm[i] * (size_t)t->dim + (size_t)d];
            if (v < mn) mn = v;
            if (v > mx) mx = v;
        }
        if (mx - mn > best_ext) { best_ext = mx - mn; best_axis = d; }
    }
    t->axis[mid] = (uint8_t)best_axis;
    if (n == 1) return;
    c.pts = t->pts; c.dim = t->dim; c.axis = best_axis;
    kd_select(t->perm + lo, n, n / 2, &c);
    kd_build(t, lo, mid);
    kd_build(t, mid + 1, hi);
}

orc_kdtree *orc_kdtree_new(const double *dst, size_t m, int dim)
{
    orc_kdtree *t = (orc_kdtree *)malloc(sizeof(*t));
    size_t i;
    t->pts = dst; t->m = m; t->dim = dim;
    t->perm = (int32_t *)malloc((m ? m : 1) * sizeof(int32_t));
    t->axis = (uint8_t *)malloc((m ? m : 1));
    for (i = 0; i < m; i++) t->perm[i] = (int32_t)i;
    kd_build(t, 0, m);
    return t;
}

void orc_kdtree_free(orc_kdtree *t)
{
    if (!t) return;
    free(t->perm);
    free(t->axis);
    free(t);
}

typedef struct { const orc_kdtree *t; const double *q; double bd; int32_t best; } kd_query;

static void kd_search(kd_query *s, size_t lo, size_t hi)
{
    const orc_kdtree *t = s->t;
    size_t mid;
    int32_t id;
    int ax;
    double d2, diff;
    if (lo >= hi) return;
    mid = lo + (hi - lo) / 2;
    id = t->perm[mid];
    d2 = dist2(t->pts + (size_t)id * (size_t)t->dim, s->q, t->dim);
    if (s->best < 0 || d2 < s->bd || (d2 == s->bd && id < s->best)) {
        s->bd = d2;
        s->best = id;
    }
    if (hi - lo == 1) return;
    ax = t->axis[mid];
    diff = s->q[ax] - t->pts[(size_t)id * (size_t)t->dim + (size_t)ax];
    if (diff < 0.) {
        kd_search(s, lo, mid);
        /* prune only when strictly farther: a tie in the far side may carry a lower index */
        if (!(diff * diff > s->bd)) kd_search(s, mid + 1, hi);
    } else {
        kd_search(s, mid + 1, hi);
        if (!(diff * diff > s->bd)) kd_search(s, lo, mid);
    }
}

int64_t orc_kdtree_search(const orc_kdtree *t, const double *q)
{
    kd_query s;
    s.t = t; s.q = q; s.bd = 0.; s.best = -1;
    kd_search(&s, 0, t->m);
    return s.best;
}

void orc_nn_search_batch(const double *dst, size_t m, int dim, const double *queries, size_t q,
                         int mode, int32_t *out_idx)
{
    size_t i;
    if (mode == ORC_NN_KDTREE) {
        orc_kdtree *t = orc_kdtree_new(dst, m, dim);
        for (i = 0; i < q; i++)
            out_idx[i] = (int32_t)orc_kdtree_search(t, queries + i * (size_t)dim);
        orc_kdtree_free(t);
    } else {
        for (i = 0; i < q; i++)
            out_idx[i] = (int32_t)orc_nn_brute(dst, m, dim, queries + i * (size_t)dim);
    }
}

/* ===================================================================================== */
uint64_t orc_mix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

/* Icp2d::estimate (src/lib.rs:105-130) and Icp3d::estimate (src/lib.rs:148-173) share this
 * skeleton; dim==3 adds transform_xy (lib.rs:52-57) and get_xy (lib.rs:86-89). */
static int icp_estimate(int dim, const double *src, size_t n, const double *dst, size_t m,
                        const double init[6], uint32_t max_iter, double out[6], orc_trace *tr)
{
    double T[6];
    double *src_t, *s2, *d2;
    orc_kdtree *tree = NULL;
    uint32_t it;
    uint64_t H = 0;
    size_t i;
    int nn_mode = tr ? tr->nn_mode : ORC_NN_BRUTE;
    memcpy(T, init, sizeof(T)); /* lib.rs:111,154 */
    memcpy(out, init, sizeof(T));
    if (m == 0 && n > 0 && max_iter > 0)
        return 1; /* reference: index.unwrap() panics (lib.rs:122,165) */
    if (tr && tr->final_cost) *tr->final_cost = 0.;
    if (tr && tr->shared_tree)
        tree = (orc_kdtree *)tr->shared_tree; /* one Icp2d::new(dst), many estimate() calls */
    else if (nn_mode == ORC_NN_KDTREE)
        tree = orc_kdtree_new(dst, m, dim); /* Icp2d::new / Icp3d::new, lib.rs:97-102,139-144 */
    src_t = (double *)malloc((n ? n : 1) * (size_t)dim * sizeof(double));
    s2 = (double *)malloc((n ? n : 1) * 2 * sizeof(double));
    d2 = (double *)malloc((n ? n : 1) * 2 * sizeof(double));
    for (it = 0; it < max_iter; it++) { /* lib.rs:112,155: no early exit */
        double dT[6], Tn[6];
        uint32_t acc, why;
        uint64_t h = 0;
        for (i = 0; i < n; i++) { /* lib.rs:113-116 / 156-159 */
            orc_transform_apply(T, src + i * (size_t)dim, src_t + i * (size_t)dim);
            if (dim == 3)
                src_t[i * 3 + 2] = src[i * 3 + 2];
        }
        for (i = 0; i < n; i++) { /* lib.rs:118-124 / 161-167 */
            const double *q = src_t + i * (size_t)dim;
            int64_t idx = tree ? orc_kdtree_search(tree, q) : orc_nn_brute(dst, m, dim, q);
            s2[2 * i] = q[0];
            s2[2 * i + 1] = q[1];
            d2[2 * i] = dst[(size_t)idx * (size_t)dim];
            d2[2 * i + 1] = dst[(size_t)idx * (size_t)dim + 1];
            h += orc_mix64(((uint64_t)i << 32) | (uint64_t)(uint32_t)idx);
            if (tr && tr->corr) {
                if (tr->corr_mode == ORC_CORR_ALL)
                    tr->corr[(size_t)it * n + i] = (int32_t)idx;
                else if (tr->corr_mode == ORC_CORR_LAST && it + 1 == max_iter)
                    tr->corr[i] = (int32_t)idx;
            }
        }
        H = H * 0x100000001B3ull + h;
        estimate_transform_ex(s2, d2, n, dT, &acc, &why, tr ? tr->margins : NULL); /* lib.rs:125,168 */
        if (tr && tr->inner)
            tr->inner[it] = acc | (why << 16);
        if (tr && tr->final_cost && it + 1 == max_iter) /* huber_error (lib.rs:45-50) of the last inner solve */
            *tr->final_cost = orc_huber_error(dT, s2, d2, n);
        orc_transform_mul(dT, T, Tn); /* lib.rs:127,170 */
        memcpy(T, Tn, sizeof(T));
    }
    memcpy(out, T, sizeof(T));
    if (tr && tr->corr_hash)
        *tr->corr_hash = H;
    free(src_t);
    free(s2);
    free(d2);
    if (!(tr && tr->shared_tree)) orc_kdtree_free(tree);
    return 0;
}

int orc_icp2d_estimate(const double *src, size_t n, const double *dst, size_t m,
                       const double init[6], uint32_t max_iter, double out[6], orc_trace *tr)
{
    return icp_estimate(2, src, n, dst, m, init, max_iter, out, tr);
}

int orc_icp3d_estimate(const double *src, size_t n, const double *dst, size_t m,
                       const double init[6], uint32_t max_iter, double out[6], orc_trace *tr)
{
    return icp_estimate(3, src, n, dst, m, init, max_iter, out, tr);
}

/* ===================================================================================== */
typedef struct {
    int dim;
    const double *src; const uint32_t *src_off;
    const double *dst; const uint32_t *dst_off;
    const double *init_T;
    uint32_t B, max_iter;
    int nn_mode;
    double *out_T; uint32_t *inner; uint64_t *corr_hash; int32_t *status;
    const orc_kdtree *shared_tree; size_t shared_m;   /* non-NULL: every pair searches this one target */
    double *final_cost, *margins;                     /* nullable [B], [B*2] */
    volatile uint32_t next;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *jb = (batch_job *)arg;
    for (;;) {
        uint32_t b = __sync_fetch_and_add(&jb->next, 1u);
        orc_trace tr;
        int st;
        if (b >= jb->B) break;
        memset(&tr, 0, sizeof(tr));
        tr.nn_mode = jb->nn_mode;
        tr.inner = jb->inner ? jb->inner + (size_t)b * jb->max_iter : NULL;
        tr.corr_hash = jb->corr_hash ? jb->corr_hash + b : NULL;
        tr.final_cost = jb->final_cost ? jb->final_cost + b : NULL;
        if (jb->margins) {
            tr.margins = jb->margins + (size_t)b * 2;
            tr.margins[0] = tr.margins[1] = HUGE_VAL;
        }
        if (jb->shared_tree || jb->dst_off == NULL) {
            tr.shared_tree = jb->shared_tree;
            st = icp_estimate(jb->dim, jb->src + (size_t)jb->src_off[b] * (size_t)jb->dim,
                              jb->src_off[b + 1] - jb->src_off[b], jb->dst, jb->shared_m,
                              jb->init_T + (size_t)b * 6, jb->max_iter, jb->out_T + (size_t)b * 6, &tr);
        } else
        st = icp_estimate(jb->dim, jb->src + (size_t)jb->src_off[b] * (size_t)jb->dim,
                          jb->src_off[b + 1] - jb->src_off[b],
                          jb->dst + (size_t)jb->dst_off[b] * (size_t)jb->dim,
                          jb->dst_off[b + 1] - jb->dst_off[b], jb->init_T + (size_t)b * 6,
                          jb->max_iter, jb->out_T + (size_t)b * 6, &tr);
        if (jb->status) jb->status[b] = st;
    }
    return NULL;
}

int orc_icp_estimate_batch(int dim, const double *src, const uint32_t *src_off,
                           const double *dst, const uint32_t *dst_off, const double *init_T,
                           uint32_t B, uint32_t max_iter, int nn_mode, int nthreads,
                           double *out_T, uint32_t *inner, uint64_t *corr_hash, int32_t *status)
{
    return orc_icp_estimate_batch_ex(dim, src, src_off, dst, dst_off, 0, init_T, B, max_iter, nn_mode,
                                     nthreads, out_T, inner, corr_hash, status, NULL, NULL);
}

int orc_icp_estimate_batch_ex(int dim, const double *src, const uint32_t *src_off,
                              const double *dst, const uint32_t *dst_off, size_t shared_m,
                              const double *init_T, uint32_t B, uint32_t max_iter, int nn_mode,
                              int nthreads, double *out_T, uint32_t *inner, uint64_t *corr_hash,
                              int32_t *status, double *final_cost, double *margins)
{
    batch_job jb;
    pthread_t *th;
    orc_kdtree *shared = NULL;
    int i, started = 0;
    jb.shared_tree = NULL; jb.shared_m = shared_m; jb.final_cost = final_cost; jb.margins = margins;
    if (dst_off == NULL && nn_mode == ORC_NN_KDTREE && shared_m > 0) {
        shared = orc_kdtree_new(dst, shared_m, dim); /* Icp2d::new once (lib.rs:97-102) */
        jb.shared_tree = shared;
    }
    if (nthreads < 1) nthreads = 1;
    if ((uint32_t)nthreads > B && B > 0) nthreads = (int)B;
    jb.dim = dim; jb.src = src; jb.src_off = src_off; jb.dst = dst; jb.dst_off = dst_off;
    jb.init_T = init_T; jb.B = B; jb.max_iter = max_iter; jb.nn_mode = nn_mode;
    jb.out_T = out_T; jb.inner = inner; jb.corr_hash = corr_hash; jb.status = status;
    jb.next = 0;
    if (inner) memset(inner, 0, (size_t)B * max_iter * sizeof(uint32_t));
    if (nthreads == 1) {
        batch_worker(&jb);
        orc_kdtree_free(shared);
        return 1;
    }
    th = (pthread_t *)malloc((size_t)nthreads * sizeof(pthread_t));
    for (i = 0; i < nthreads; i++)
        if (pthread_create(&th[started], NULL, batch_worker, &jb) == 0) started++;
    if (started == 0) batch_worker(&jb);
    for (i = 0; i < started; i++) pthread_join(th[i], NULL);
    free(th);
    orc_kdtree_free(shared);
    return started ? started : 1;
}
