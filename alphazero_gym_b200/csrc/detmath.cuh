// detmath.cuh -- deterministic elementary functions for the engine.
//
// Every function is a fixed sequence of correctly rounded IEEE-754 operations (add, mul, fma, div,
// sqrt, rint), so its result is a pure function of the input bits on any conforming machine.  The spec
// (Cephes expf/tanhf polynomials, fdlibm __kernel_sin/__kernel_cos/log, two-step
// Cody-Waite reductions) is what makes engine results reproducible bit for bit by an independent CPU
// implementation.  This translation unit is compiled with -fmad=false: a*b+c is never contracted, fused
// operations appear only where __fmaf_rn / __fma_rn is written.
//
// Why not libdevice: CUDA's sin/cos/expf are accurate to 1-2 ulp but are not the same function as
// glibc's or torch's, so "bit-exact visit counts" could not be checked against anything.
#pragma once
#include <math_constants.h>

namespace det {

// exp(r) - 1 on |r| <= ln2/2:  r + r^2 P(r)
__device__ __forceinline__ float expm1_poly(float r) {
    float p = 1.9875691500E-4f;
    p = __fmaf_rn(p, r, 1.3981999507E-3f);
    p = __fmaf_rn(p, r, 8.3334519073E-3f);
    p = __fmaf_rn(p, r, 4.1665795894E-2f);
    p = __fmaf_rn(p, r, 1.6666665459E-1f);
    p = __fmaf_rn(p, r, 5.0000001201E-1f);
    const float z = __fmul_rn(r, r);
    return __fmaf_rn(p, z, r);
}

__device__ __forceinline__ float exp_reduce(float x, float& n) {
    n = rintf(__fmul_rn(x, 1.44269504088896341f));
    float r = __fmaf_rn(n, -0.693359375f, x);
    r = __fmaf_rn(n, 2.12194440e-4f, r);
    return r;
}

__device__ __forceinline__ float expf_(float x) {
    if (!(x <= 88.0f)) return (x != x) ? x : CUDART_INF_F;
    if (x < -87.0f) return 0.0f;
    float n;
    const float r = exp_reduce(x, n);
    const float y = __fadd_rn(expm1_poly(r), 1.0f);
    return __uint_as_float(__float_as_uint(y) + ((uint32_t)(int32_t)n << 23));
}

__device__ __forceinline__ float expm1f_(float x) {
    if (x != x) return x;
    if (x < -17.5f) return -1.0f;
    if (x > 88.0f) return CUDART_INF_F;
    float n;
    const float r = exp_reduce(x, n);
    const float p = expm1_poly(r);
    if (n == 0.0f) return p;
    const float t = __uint_as_float((uint32_t)((int32_t)n + 127) << 23);
    return __fmaf_rn(p, t, __fsub_rn(t, 1.0f));
}

__device__ __forceinline__ float tanhf_(float x) {
    const float ax = fabsf(x);
    float y;
    if (ax != ax) return x;
    if (ax >= 9.1f) {
        y = 1.0f;
    } else if (ax >= 0.625f) {
        const float e = expf_(__fadd_rn(ax, ax));
        y = __fsub_rn(1.0f, __fdiv_rn(2.0f, __fadd_rn(e, 1.0f)));
    } else {
        const float z = __fmul_rn(ax, ax);
        float p = -5.70498872745E-3f;
        p = __fmaf_rn(p, z, 2.06390887954E-2f);
        p = __fmaf_rn(p, z, -5.37397155531E-2f);
        p = __fmaf_rn(p, z, 1.33314422036E-1f);
        p = __fmaf_rn(p, z, -3.33332819422E-1f);
        y = __fmaf_rn(__fmul_rn(p, z), ax, ax);
    }
    return copysignf(y, x);
}

// ---- double precision (plain operators: this file is built with -fmad=false) ---------------------
__device__ __forceinline__ double ksin(double x, double y) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03, S3 = -1.98412698298579493134e-04,
                 S4 = 2.75573137070700676789e-06, S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    const double z = x * x;
    const double w = z * z;
    const double r = (S2 + z * (S3 + z * S4)) + (z * w) * (S5 + z * S6);
    const double v = z * x;
    return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

__device__ __forceinline__ double kcos(double x, double y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03, C3 = 2.48015872894767294178e-05,
                 C4 = -2.75573143513906633035e-07, C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    const double z = x * x;
    const double w = z * z;
    const double r = z * (C1 + z * (C2 + z * C3)) + (w * w) * (C4 + z * (C5 + z * C6));
    const double hz = 0.5 * z;
    const double w1 = 1.0 - hz;
    return w1 + (((1.0 - w1) - hz) + (z * r - x * y));
}

__device__ __forceinline__ int rem_pio2(double x, double& y0, double& y1) {
    const double invpio2 = 6.36619772367581382433e-01, pio2_1 = 1.57079632673412561417e+00,
                 pio2_2 = 6.07710050630396597660e-11, pio2_2t = 2.02226624879595063154e-21;
    const double fn = rint(x * invpio2);
    const double t = x - fn * pio2_1;
    double w = fn * pio2_2;
    const double r = t - w;
    w = fn * pio2_2t - ((t - r) - w);
    y0 = r - w;
    y1 = (r - y0) - w;
    return (int)fn;
}

__device__ __forceinline__ void sincos_(double x, double& s, double& c) {
    double y0, y1;
    const int n = rem_pio2(x, y0, y1);
    const double ks = ksin(y0, y1), kc = kcos(y0, y1);
    switch (n & 3) {
        case 0: s = ks; c = kc; break;
        case 1: s = kc; c = -ks; break;
        case 2: s = -ks; c = -kc; break;
        default: s = -kc; c = ks; break;
    }
}

__device__ __forceinline__ double sin_(double x) { double s, c; sincos_(x, s, c); return s; }
__device__ __forceinline__ double cos_(double x) { double s, c; sincos_(x, s, c); return c; }

// log(x), x positive normal
__device__ __forceinline__ double log_(double x) {
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    uint32_t hx = (uint32_t)__double2hiint(x);
    const uint32_t lx = (uint32_t)__double2loint(x);
    hx += 0x3ff00000u - 0x3fe6a09eu;
    const int k = (int)(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
    x = __hiloint2double((int)hx, (int)lx);
    const double f = x - 1.0;
    const double hfsq = 0.5 * f * f;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    const double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    const double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    const double R = t2 + t1;
    const double dk = (double)k;
    return ((((s * (hfsq + R)) + dk * ln2_lo) - hfsq) + f) + dk * ln2_hi;
}

}  // namespace det
