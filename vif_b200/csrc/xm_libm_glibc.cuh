// glibc 2.39's double sin and cos for |x| < 2.426265, restated bit for bit (host and device).
//
// The reference computes cos(dec) and sin(half differences) with the host's libm (qxmatch.hpp:174-180,
// 192-204); its proxies -- and therefore the order of near-tied candidates -- depend on that libm's last
// bits. glibc is the one third-party dependency on the path; it is not part of
// /root/reference, so its PUBLISHED algorithm (sysdeps/ieee754/dbl-64/s_sin.c, glibc 2.39, the version
// of this image: `ldd --version`) is restated here:
//   |x| < 2^-26 (sin) / 2^-27 (cos): x / 1
//   |x| < 0.855469:   do_sin / do_cos: |x| is split at a multiple of 2^-7 (u = big + |x|), the tabulated
//                     sin/cos of that multiple (double-double, xm_sincostab.h) is combined with short
//                     polynomials of the remainder; do_sin takes a Taylor branch below 0.126
//   |x| < 2.426265:   sin(x) = do_cos(pi/2 - |x|), cos(x) = do_sin(pi/2 - |x|), pi/2 as hp0 + hp1
// Every operation is plain IEEE double arithmetic in a fixed order, so the only freedom is
// contraction: x86-64 glibc carries an -mfma build of the same source that it selects at run time on
// FMA hosts. MODEL 1 contracts every a*b+c the way gcc does for that build, MODEL 2 never contracts.
// xm_init() probes the host libm and picks the model (xm_geometry.cpp); both were checked against
// __sin_fma/__cos_fma and __sin_sse2/__cos_sse2 of libm-2.39.a on 4 x 2e7 random arguments: 0 mismatches.
// Beyond 2.426265 (half differences of sources more than 278 deg apart in RA) the CUDA libm is used.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "xm_sincostab.h"

#ifdef __CUDACC__
#define XM_LM_HD __host__ __device__ __forceinline__
#else
#define XM_LM_HD inline
#endif

namespace xm {
namespace glibc {

static const double h_sincostab[XM_SINCOSTAB_SIZE] = XM_SINCOSTAB_VALUES;
#ifdef __CUDACC__
static __device__ const double d_sincostab[XM_SINCOSTAB_SIZE] = XM_SINCOSTAB_VALUES;
#endif

XM_LM_HD const double* table() {
#ifdef __CUDA_ARCH__
    return d_sincostab;
#else
    return h_sincostab;
#endif
}

// separately rounded IEEE operations whatever the compiler's contraction setting
XM_LM_HD double mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a*b; return r;
#endif
}
XM_LM_HD double add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    volatile double r = a + b; return r;
#endif
}
// a*b + c: fused in the FMA build of glibc (model 1), two roundings otherwise (model 2)
XM_LM_HD double F(int model, double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return model == 1 ? __fma_rn(a, b, c) : __dadd_rn(__dmul_rn(a, b), c);
#else
    return model == 1 ? fma(a, b, c) : add(mul(a, b), c);
#endif
}

constexpr double kBig = 0x1.8p45;
constexpr double kSn3 = -1.66666666666664880952546298448555E-01, kSn5 = 8.33333214285722277379541354343671E-03;
constexpr double kCs2 = 4.99999999999999999999950396842453E-01, kCs4 = -4.16666666666664434524222570944589E-02,
                 kCs6 = 1.38888874007937613028114285595617E-03;
constexpr double kHp0 = 1.5707963267948966, kHp1 = 6.123233995736766e-17;
constexpr double kS1 = -0x1.5555555555555p-3, kS2 = 0x1.1111111110ECEp-7, kS3 = -0x1.A01A019DB08B8p-13,
                 kS4 = 0x1.71DE27B9A7ED9p-19, kS5 = -0x1.ADDFFC2FCDF59p-26;

XM_LM_HD uint32_t hi_word(double x) {
    uint64_t b;
    memcpy(&b, &x, 8);
    return (uint32_t)(b >> 32);
}
XM_LM_HD int table_index(double ux) {           // low word of big + |x| = |x| rounded to 2^-7, times 4
    uint64_t b;
    memcpy(&b, &ux, 8);
    return (int)(uint32_t)b*4;
}

XM_LM_HD double do_cos(int m, double x, double dx) {
    if (x < 0) dx = -dx;
    const double ux = add(kBig, fabs(x));
    const int k = table_index(ux);
    x = add(add(fabs(x), -add(ux, -kBig)), dx);
    const double xx = mul(x, x);
    const double s = F(m, mul(x, xx), F(m, xx, kSn5, kSn3), x);
    const double c = mul(xx, F(m, xx, F(m, xx, kCs6, kCs4), kCs2));
    const double* t = table();
    const double sn = t[k], ssn = t[k + 1], cs = t[k + 2], ccs = t[k + 3];
    const double cor = F(m, -sn, s, F(m, -cs, c, F(m, -s, ssn, ccs)));
    return add(cs, cor);
}

XM_LM_HD double taylor_sin(int m, double xx, double x, double dx) {
    const double p = F(m, F(m, F(m, F(m, kS5, xx, kS4), xx, kS3), xx, kS2), xx, kS1);
    const double t = F(m, F(m, p, x, -mul(0.5, dx)), xx, dx);
    return add(x, t);
}

XM_LM_HD double do_sin(int m, double x, double dx) {
    const double xold = x;
    if (fabs(x) < 0.126) return taylor_sin(m, mul(x, x), x, dx);
    if (x <= 0) dx = -dx;
    const double ux = add(kBig, fabs(x));
    const int k = table_index(ux);
    x = add(fabs(x), -add(ux, -kBig));
    const double xx = mul(x, x);
    const double s = add(x, F(m, mul(x, xx), F(m, xx, kSn5, kSn3), dx));
    const double c = F(m, x, dx, mul(xx, F(m, xx, F(m, xx, kCs6, kCs4), kCs2)));
    const double* t = table();
    const double sn = t[k], ssn = t[k + 1], cs = t[k + 2], ccs = t[k + 3];
    const double cor = F(m, cs, s, F(m, -sn, c, F(m, s, ccs, ssn)));
    return copysign(add(sn, cor), xold);
}

// true: *out = glibc's cos(x) / sin(x); false: |x| outside the restated range
XM_LM_HD bool cos_small(int m, double x, double* out) {
    const uint32_t k = hi_word(x) & 0x7fffffffu;
    if (k < 0x3e400000u) { *out = 1.0; return true; }
    if (k < 0x3feb6000u) { *out = do_cos(m, x, 0.0); return true; }
    if (k < 0x400368fdu) {
        const double y = add(kHp0, -fabs(x));
        const double a = add(y, kHp1);
        const double da = add(add(y, -a), kHp1);
        *out = do_sin(m, a, da);
        return true;
    }
    return false;
}
XM_LM_HD bool sin_small(int m, double x, double* out) {
    const uint32_t k = hi_word(x) & 0x7fffffffu;
    if (k < 0x3e500000u) { *out = x; return true; }
    if (k < 0x3feb6000u) { *out = do_sin(m, x, 0.0); return true; }
    if (k < 0x400368fdu) {
        const double t = add(kHp0, -fabs(x));
        *out = copysign(do_cos(m, t, kHp1), x);
        return true;
    }
    return false;
}

}  // namespace glibc
}  // namespace xm
