// physmath.cuh -- per-particle fp64 formulas shared by the pair-parallel (MEX path) and the CSR gather
// (device-resident) kernels.
#pragma once
#include <cuda_runtime.h>

namespace sphb {

// Reciprocal and reciprocal square root of a normal positive double: the SFU seed (MUFU.RCP64H /
// MUFU.RSQ64H, relative error d <= 2^-22) refined on the FP64 pipe to ~1 ulp.  Used only by the physics
// gathers (1e-11 gate); every decision about list membership uses IEEE sqrt and divide (neighbor.cu).
//
// SPHB_NEWTON_CUBIC=1: ONE third-order step instead of two second-order ones.  With e = 1 - x y (= -d) the
// reciprocal is y / (1 - e) = y (1 + e + e^2 + O(e^3)); with e = 1 - x y^2 (= -2 d) the reciprocal square root is
// y / sqrt(1 - e) = y (1 + e / 2 + 3 e^2 / 8 + O(e^3)): the truncated terms are below 2^-64.  3 instead of 4 and
// 5 instead of 8 FP64 instructions, and two links fewer in the dependent chain every pair starts with.
#ifndef SPHB_NEWTON_CUBIC
#define SPHB_NEWTON_CUBIC 1
#endif
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if SPHB_NEWTON_CUBIC
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
#else
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    return fma(y, e, y);
#endif
}
__device__ __forceinline__ double fast_rsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#if SPHB_NEWTON_CUBIC
    const double e = fma(-(x * y), y, 1.0);
    return fma(y, e * fma(0.375, e, 0.5), y);
#else
    double e = fma(-x * y, y, 1.0);
    y = fma(y * 0.5, e, y);
    e = fma(-x * y, y, 1.0);
    return fma(y * 0.5, e, y);
#endif
}

// KGC matrix from the accumulated A (ph:321-366): regularised pseudo-inverse blended with I.
__device__ __forceinline__ void kgc_from_A(double a11, double a12, double a21, double a22, double &b11, double &b12, double &b21,
                           double &b22)
{
    const double eps_reg = 1e-8;
    const double m11 = a11 * a11 + a21 * a21 + eps_reg;
    const double m12 = a11 * a12 + a21 * a22;
    const double m22 = a12 * a12 + a22 * a22 + eps_reg;
    const double det_m = m11 * m22 - m12 * m12;
    double p11 = 1.0, p12 = 0.0, p21 = 0.0, p22 = 1.0;
    if (!(fabs(det_m) < 1e-20)) {
        const double i11 = m22 / det_m, i12 = -m12 / det_m, i22 = m11 / det_m;
        p11 = i11 * a11 + i12 * a12;
        p12 = i11 * a21 + i12 * a22;
        p21 = i12 * a11 + i22 * a12;
        p22 = i12 * a21 + i22 * a22;
    }
    const double det_a = a11 * a22 - a12 * a21;
    const double lack = fmax(1.0 - det_a, 0.0);
    const double denom = det_a + lack;
    double w1 = 0.0, w2 = 1.0;
    if (!(fabs(denom) < 1e-12)) {
        w1 = det_a / denom;
        w2 = lack / denom;
    }
    b11 = w1 * p11 + w2;
    b12 = w1 * p12;
    b21 = w1 * p21;
    b22 = w1 * p22 + w2;
}

}  // namespace sphb
