// Reference-order FP64 arithmetic for the parity-critical expressions.
//
// Every operation here is spelled with __d*_rn / __fma_rn intrinsics so that nvcc can neither
// contract nor reassociate, whatever -fmad says: x86-64 gcc without -mfma (the reference's flags,
// CMakeLists.txt:54-58) evaluates the same expressions as separately rounded IEEE operations.
// IEEE + - * / sqrt floor are then bit-identical host/device; only sin/cos/asin can differ.
//
// sin(): glibc 2.39's double sin takes, for 2^-26 <= |x| < 0.126, a degree-11 odd Taylor branch
// (sysdeps/ieee754/dbl-64/s_sin.c, TAYLOR_SIN with dx = 0). That branch is plain IEEE arithmetic,
// so it is restated here and is bit-identical to the host libm on the whole range the cell-binned
// search ever feeds it (half-differences of coordinates a few cells apart). On x86-64 hosts with
// FMA, glibc dispatches to the -mfma build of the same source, in which the Horner steps contract;
// xm_init() probes the host libm on a dozen discriminating inputs and selects the matching model
// (GridView::sin_model). Outside that range the CUDA libm sin (<= 2 ulp) is used.
#pragma once
#include "xm_common.cuh"
#include "xm_libm_glibc.cuh"

namespace xm {

__device__ __forceinline__ double sin_taylor_glibc(double x, int model) {
    const double s1 = -0x1.5555555555555p-3, s2 = 0x1.1111111110ECEp-7, s3 = -0x1.A01A019DB08B8p-13,
                 s4 = 0x1.71DE27B9A7ED9p-19, s5 = -0x1.ADDFFC2FCDF59p-26;
    const double xx = __dmul_rn(x, x);
    double p;
    if (model == 1) {
        p = __fma_rn(__fma_rn(__fma_rn(__fma_rn(s5, xx, s4), xx, s3), xx, s2), xx, s1);
    } else {
        p = __dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(
                __dmul_rn(s5, xx), s4), xx), s3), xx), s2), xx), s1);
    }
    const double t = __dmul_rn(__dmul_rn(p, x), xx);
    return __dadd_rn(x, t);
}

__device__ __forceinline__ double sin_ref(double x, int model) {
    if (model == 0) return sin(x);          // host libm not recognised (xm_init): CUDA libm, <= 2 ulp
    const double ax = fabs(x);
    if (ax < 0.126) {
        if (ax < 0x1p-26) return x;
        return sin_taylor_glibc(x, model);
    }
    // 0.126 <= |x| < 2.426: glibc's table branches restated (xm_libm_glibc.cuh); beyond: CUDA libm (<= 2 ulp)
    double r;
    if (glibc::sin_small(model, x, &r)) return r;
    return sin(x);
}

// cos(dec) of a catalog row as the host libm computes it (qxmatch.hpp:176,180). model = the libm model
// xm_init() recognised (1 FMA build, 2 plain build of glibc 2.39); 0 = not recognised: CUDA libm, and the
// exact kernels' ambiguity counter (xm_stats.rows_ambiguous) stays the only guard.
__device__ __forceinline__ double cos_ref(double x, int model) {
    double r;
    if (model != 0 && glibc::cos_small(model, x, &r)) return r;
    return cos(x);
}

// same, for callers that guarantee |x| < 0.126 (coordinates a few cells apart): no libm fallback
__device__ __forceinline__ double sin_ref_small(double x, int model) {
    if (model == 0) return sin(x);
    if (fabs(x) < 0x1p-26) return x;
    return sin_taylor_glibc(x, model);
}

__device__ __forceinline__ double proxy_sph_small(double x1, double y1, double c1, double x2,
                                                  double y2, double c2, int model) {
    const double sra = sin_ref_small(__dmul_rn(0.5, __dsub_rn(x2, x1)), model);
    const double sde = sin_ref_small(__dmul_rn(0.5, __dsub_rn(y2, y1)), model);
    return __dadd_rn(__dmul_rn(sde, sde),
                     __dmul_rn(__dmul_rn(__dmul_rn(sra, sra), c2), c1));
}

// qxmatch.hpp:200-202: sde*sde + sra*sra*dcdec2*dcdec1, left to right.
__device__ __forceinline__ double proxy_sph(double x1, double y1, double c1, double x2, double y2,
                                            double c2, int model) {
    const double sra = sin_ref(__dmul_rn(0.5, __dsub_rn(x2, x1)), model);
    const double sde = sin_ref(__dmul_rn(0.5, __dsub_rn(y2, y1)), model);
    return __dadd_rn(__dmul_rn(sde, sde),
                     __dmul_rn(__dmul_rn(__dmul_rn(sra, sra), c2), c1));
}

// qxmatch.hpp:198: sqr(ra2 - ra1) + sqr(dec2 - dec1) on the raw inputs.
__device__ __forceinline__ double proxy_lin(double x1, double y1, double x2, double y2) {
    const double dx = __dsub_rn(x2, x1), dy = __dsub_rn(y2, y1);
    return __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
}

// qxmatch.hpp:206-214
__device__ __forceinline__ double true_to_proxy(double dist, bool linear, int model) {
    if (linear) return __dmul_rn(dist, dist);
    const double k = __dmul_rn(__dmul_rn(3600.0, __ddiv_rn(180.0, XM_DPI)), 2.0);
    const double s = sin_ref(__ddiv_rn(dist, k), model);
    return __dmul_rn(s, s);
}

// qxmatch.hpp:215-222: dist = 3600*(180/dpi)*2*asin(sqrt(p)).
// For s = sqrt(p) < 0.01 (sources closer than ~1.1 deg: every binned match)
// asin(s) = s(1 + s^2/6 + 3 s^4/40 + 15 s^6/336 + 105 s^8/3456 + ...), truncation error after the s^8 term
// below 3e-22 relative -- 9 FP64 operations instead of the libm call, within 1 ulp of the correctly
// rounded result (north-star tolerance on distances: 1e-9 relative). EVERY kernel converts through this
// one function, so a distance does not depend on which kernel (filter or exact) decided its row.
__device__ __forceinline__ double proxy_to_true(double p, bool linear) {
    if (linear) return __dsqrt_rn(p);
    const double k = __dmul_rn(__dmul_rn(3600.0, __ddiv_rn(180.0, XM_DPI)), 2.0);
    const double s = __dsqrt_rn(p);
    if (!(s < 0.01)) return __dmul_rn(k, asin(s));
    const double t = __fma_rn(p, __fma_rn(p, __fma_rn(p, 105.0/3456.0, 15.0/336.0), 3.0/40.0), 1.0/6.0);
    return __dmul_rn(k, __fma_rn(__dmul_rn(s, p), t, s));
}

__device__ __forceinline__ double proxy_to_true_small(double p) { return proxy_to_true(p, false); }

// astro.hpp:33-39 (angdist, degrees -> arcsec) with the first point already converted:
// x1r = d2r*ra1, y1r = d2r*dec1, c1 = cos(y1r).
__device__ __forceinline__ double angdist_pre(double x1r, double y1r, double c1, double tra2,
                                              double tdec2, int model) {
    const double d2r = XM_D2R;
    const double ra2 = __dmul_rn(d2r, tra2), dec2 = __dmul_rn(d2r, tdec2);
    const double sra = sin_ref(__dmul_rn(0.5, __dsub_rn(ra2, x1r)), model);
    const double sde = sin_ref(__dmul_rn(0.5, __dsub_rn(dec2, y1r)), model);
    const double h = __dadd_rn(__dmul_rn(sde, sde),
                               __dmul_rn(__dmul_rn(__dmul_rn(sra, sra), cos_ref(dec2, model)), c1));
    return __ddiv_rn(__dmul_rn(__dmul_rn(3600.0, 2.0), asin(__dsqrt_rn(h))), d2r);
}

}  // namespace xm
