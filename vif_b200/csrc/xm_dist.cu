// Distance helpers of the same haversine family as the cross-match:
//   angdist, vector forms       astro/astro.hpp:33-81     [degrees in, arcsec out]
//   angdist_less                astro/astro.hpp:83-103    proxy < sqr(sin(radius/2)), no asin
//   qdist                       astro/qxmatch.hpp:684-717 n x n matrix, ret(j,i) for j > i, 0 elsewhere
// Reference operation order throughout (xm_math_exact.cuh). The sine is bit-identical to glibc for
// half-differences below 0.126 rad and CUDA's (<= 2 ulp) beyond; cos / asin are CUDA's: distances
// agree with the reference to ~1e-15 relative, booleans except within that band of the radius.
#include "xm_common.cuh"
#include "xm_math_exact.cuh"

namespace xm {

namespace {

// astro.hpp:33-39
__device__ __forceinline__ double angdist_deg(double tra1, double tdec1, double tra2, double tdec2, int model) {
    const double d2r = XM_D2R;
    const double ra1 = __dmul_rn(d2r, tra1), ra2 = __dmul_rn(d2r, tra2);
    const double dec1 = __dmul_rn(d2r, tdec1), dec2 = __dmul_rn(d2r, tdec2);
    const double sra = sin_ref(__dmul_rn(0.5, __dsub_rn(ra2, ra1)), model);
    const double sde = sin_ref(__dmul_rn(0.5, __dsub_rn(dec2, dec1)), model);
    const double h = __dadd_rn(__dmul_rn(sde, sde),
                               __dmul_rn(__dmul_rn(__dmul_rn(sra, sra), cos_ref(dec2, model)), cos_ref(dec1, model)));
    return __ddiv_rn(__dmul_rn(__dmul_rn(3600.0, 2.0), asin(__dsqrt_rn(h))), d2r);
}

__global__ void __launch_bounds__(256)
angdist_kernel(const double* __restrict__ ra1, const double* __restrict__ dec1, const double* __restrict__ ra2,
               const double* __restrict__ dec2, uint64_t n, int scalar2, int model, double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t k = scalar2 ? 0 : i;
    out[i] = angdist_deg(ra1[i], dec1[i], ra2[k], dec2[k], model);
}

// astro.hpp:92-100 with the loop invariants (ra2, dec2 in radians, cos(dec2), crad) hoisted to the host
__global__ void __launch_bounds__(256)
angdist_less_kernel(const double* __restrict__ tra1, const double* __restrict__ tdec1, uint64_t n, double ra2,
                    double dec2, double cdec2, double crad, int model, uint8_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d2r = XM_D2R;
    const double ra1 = __dmul_rn(d2r, tra1[i]), dec1 = __dmul_rn(d2r, tdec1[i]);
    const double sra = sin_ref(__dmul_rn(0.5, __dsub_rn(ra2, ra1)), model);
    const double sde = sin_ref(__dmul_rn(0.5, __dsub_rn(dec2, dec1)), model);
    const double p = __dadd_rn(__dmul_rn(sde, sde), __dmul_rn(__dmul_rn(__dmul_rn(sra, sra), cdec2), cos_ref(dec1, model)));
    out[i] = p < crad ? 1 : 0;
}

// test hook: the device side of the glibc restatement
__global__ void __launch_bounds__(256)
libm_eval_kernel(int which, int model, const double* __restrict__ x, uint64_t n, double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r;
    const bool ok = which == 0 ? glibc::sin_small(model, x[i], &r) : glibc::cos_small(model, x[i], &r);
    out[i] = ok ? r : __longlong_as_double(0x7ff8000000000000ll);
}

// qxmatch.hpp:690-693
__global__ void __launch_bounds__(256)
qdist_prep_kernel(const double* __restrict__ ra, const double* __restrict__ dec, uint64_t n,
                  int model, double* __restrict__ x, double* __restrict__ y, double* __restrict__ c) {
    const uint64_t i = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double d2r = XM_D2R;
    x[i] = __dmul_rn(ra[i], d2r);
    y[i] = __dmul_rn(dec[i], d2r);
    c[i] = cos_ref(y[i], model);
}

// row j of the matrix per blockIdx.y (grid-stride), columns i along threads: coalesced row stores
__global__ void __launch_bounds__(256)
qdist_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ c, uint64_t n,
             int model, double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double xi = x[i], yi = y[i], ci = c[i];
    for (uint64_t j = blockIdx.y; j < n; j += gridDim.y) {
        double v = 0.0;
        if (j > i) v = proxy_to_true(proxy_sph(xi, yi, ci, x[j], y[j], c[j], model), false);     // :698-703
        out[j*n + i] = v;
    }
}

}  // namespace

int32_t launch_angdist(const double* ra1, const double* dec1, const double* ra2, const double* dec2, uint64_t n,
                       bool scalar2, int model, double* out, cudaStream_t s, Error& err) {
    if (n == 0) return XM_OK;
    angdist_kernel<<<(unsigned)((n + 255)/256), 256, 0, s>>>(ra1, dec1, ra2, dec2, n, scalar2 ? 1 : 0, model, out);
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

int32_t launch_angdist_less(const double* ra1, const double* dec1, uint64_t n, double tra2, double tdec2,
                            double radius, int model, uint8_t* out, cudaStream_t s, Error& err) {
    if (n == 0) return XM_OK;
    // astro.hpp:88-90, 93-94 on the host (host libm, as the reference)
    const double d2r = XM_D2R;
    const double rrad = d2r*radius/3600.0;
    const double sr = sin(rrad/2.0);
    const double crad = sr*sr;
    const double ra2 = d2r*tra2, dec2 = d2r*tdec2;
    angdist_less_kernel<<<(unsigned)((n + 255)/256), 256, 0, s>>>(ra1, dec1, n, ra2, dec2, cos(dec2), crad, model, out);
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

size_t qdist_scratch_bytes(uint64_t n) { return (size_t)n*3*sizeof(double); }

int32_t launch_qdist(const double* ra, const double* dec, uint64_t n, int model, void* scratch, double* out,
                     cudaStream_t s, Error& err) {
    if (n == 0) return XM_OK;
    double* x = (double*)scratch;
    double* y = x + n;
    double* c = y + n;
    const unsigned bx = (unsigned)((n + 255)/256);
    qdist_prep_kernel<<<bx, 256, 0, s>>>(ra, dec, n, model, x, y, c);
    const dim3 grid(bx, (unsigned)(n < 32768 ? n : 32768));
    qdist_kernel<<<grid, 256, 0, s>>>(x, y, c, n, model, out);
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

int32_t launch_libm_eval(int which, int model, const double* x, uint64_t n, double* out, cudaStream_t s, Error& err) {
    if (n == 0) return XM_OK;
    libm_eval_kernel<<<(unsigned)((n + 255)/256), 256, 0, s>>>(which, model, x, n, out);
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

}  // namespace xm
