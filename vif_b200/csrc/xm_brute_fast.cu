// brute_force path (qxmatch.hpp:490-526) as a filter + exact-refine pair, FP64-pipe bound.
//
// Filter metric: with unit vectors u = (cos d cos a, cos d sin a, sin d) the haversine proxy is
// exactly p = (1 - u1.u2)/2, so the nth nearest candidates are those with the LARGEST dot products:
// 1 DMUL + 2 DFMA + 1 DSETP per pair instead of two libm sines. The computed dot differs from the
// reference's computed proxy by an ABSOLUTE error E <= 3e-15 in p (component errors of the unit
// vectors <= 1e-15, three FMAs, and the reference's own rounding; derivation in DESIGN.md section 4).
//   * a row's top-K order is certain when all adjacent gaps among its K+1 largest dots exceed
//     kBand = 4e-14 (> 4E, in dot units); its ids are then final and its distances are computed
//     from the reference-order proxy of the winning pairs;
//   * any other row is handed to brute_refine_kernel: one CTA re-scans the candidates, evaluates
//     the reference-order proxy only for those within kBand of the row's K-th dot (nobody else can
//     enter the exact top K) and keeps the K smallest by (proxy, j) -- ascending j is the
//     reference's arrival order (qxmatch.hpp:519-523), strict '<' keeps the earlier j on ties.
// Candidate tiles are staged in shared memory by the TMA engine (cp.async.bulk + mbarrier, double
// buffered); every thread reads the same candidate at the same time (bank-conflict-free broadcast)
// and owns R sources in registers.
#include "xm_common.cuh"
#include "xm_math_exact.cuh"

namespace xm {

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 512;                  // candidates per shared-memory stage (16 KB)
constexpr double kBand = 4e-14;
constexpr uint32_t kNone = 0xffffffffu;

struct __align__(32) UVec { double ux, uy, uz, pad; };

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000ll); }

__global__ void unit_vectors_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                    const double* __restrict__ c, uint32_t n, UVec* __restrict__ out) {
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, co;
    sincos(x[i], &s, &co);
    UVec u;
    u.ux = c[i]*co; u.uy = c[i]*s; u.uz = sin(y[i]); u.pad = 0.0;
    out[i] = u;
}

// ---- TMA bulk copy helpers (1-D, no tensor map) ----------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

// descending list of the K+1 largest dots of one source
template <int K>
struct TopDots {
    double d[K + 1];
    uint32_t j[K + 1];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k <= K; ++k) { d[k] = -dinf(); j[k] = kNone; }
    }
    __device__ __forceinline__ void insert(double dot, uint32_t cj) {     // caller checked dot > d[K]
        d[K] = dot; j[K] = cj;
#pragma unroll
        for (int k = K - 1; k >= 0; --k) {
            if (d[k] < d[k + 1]) {
                const double td = d[k]; d[k] = d[k + 1]; d[k + 1] = td;
                const uint32_t tj = j[k]; j[k] = j[k + 1]; j[k + 1] = tj;
            }
        }
    }
};

template <int K, int R, bool SELF>
__global__ void __launch_bounds__(kThreads)
brute_filter_kernel(int sin_model, const UVec* __restrict__ u1, uint32_t n1, const UVec* __restrict__ u2,
                    uint32_t n2, const double* __restrict__ x1, const double* __restrict__ y1,
                    const double* __restrict__ c1, const double* __restrict__ x2,
                    const double* __restrict__ y2, const double* __restrict__ c2,
                    uint64_t plane_stride, uint64_t* __restrict__ id_out, double* __restrict__ d_out,
                    uint32_t* __restrict__ refine_list, double* __restrict__ refine_thr,
                    unsigned long long* refine_count) {
    __shared__ UVec stage[2][kTile];
    __shared__ __align__(8) uint64_t bar[2];

    const uint32_t row0 = blockIdx.x*(kThreads*R);
    double ax[R], ay[R], az[R];
    uint32_t row[R];
    TopDots<K> top[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        row[r] = row0 + r*kThreads + threadIdx.x;
        const bool live = row[r] < n1;
        const UVec u = live ? u1[row[r]] : UVec{0.0, 0.0, 0.0, 0.0};
        ax[r] = u.ux; ay[r] = u.uy; az[r] = u.uz;
        top[r].init();
    }

    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint32_t ntiles = (n2 + kTile - 1)/kTile;
    if (threadIdx.x == 0) {
        const uint32_t cnt = n2 < (uint32_t)kTile ? n2 : (uint32_t)kTile;
        mbar_expect_tx(&bar[0], cnt*(uint32_t)sizeof(UVec));
        bulk_g2s(&stage[0][0], u2, cnt*(uint32_t)sizeof(UVec), &bar[0]);
    }
    for (uint32_t t = 0; t < ntiles; ++t) {
        const uint32_t s = t & 1u;
        if (threadIdx.x == 0 && t + 1 < ntiles) {
            // stage s^1 was last read during tile t-1; the __syncthreads closing that tile ordered it
            const uint32_t base = (t + 1)*kTile;
            const uint32_t cnt = (n2 - base) < (uint32_t)kTile ? (n2 - base) : (uint32_t)kTile;
            mbar_expect_tx(&bar[s ^ 1u], cnt*(uint32_t)sizeof(UVec));
            bulk_g2s(&stage[s ^ 1u][0], u2 + base, cnt*(uint32_t)sizeof(UVec), &bar[s ^ 1u]);
        }
        mbar_wait(&bar[s], (t >> 1) & 1u);
        const uint32_t base = t*kTile;
        const uint32_t cnt = (n2 - base) < (uint32_t)kTile ? (n2 - base) : (uint32_t)kTile;
        // a tile holds some source's own row only if it overlaps this CTA's rows
        const bool self_tile = SELF && base < row0 + kThreads*R && base + cnt > row0;
        const UVec* __restrict__ st = stage[s];
        if (!self_tile) {
#pragma unroll 4
            for (uint32_t jj = 0; jj < cnt; ++jj) {
                const double cx = st[jj].ux, cy = st[jj].uy, cz = st[jj].uz;
                double dot[R];
                bool any = false;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    dot[r] = fma(ax[r], cx, fma(ay[r], cy, az[r]*cz));
                    any |= dot[r] > top[r].d[K];
                }
                if (any) {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (dot[r] > top[r].d[K]) top[r].insert(dot[r], base + jj);
                }
            }
        } else {
            for (uint32_t jj = 0; jj < cnt; ++jj) {
                const double cx = st[jj].ux, cy = st[jj].uy, cz = st[jj].uz;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const double dot = fma(ax[r], cx, fma(ay[r], cy, az[r]*cz));
                    if (base + jj != row[r] && dot > top[r].d[K]) top[r].insert(dot, base + jj);
                }
            }
        }
        __syncthreads();
    }

    // ---- epilogue: certain rows are final, the others go to the refine kernel ---------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (row[r] >= n1) continue;
        const uint32_t i = row[r];
        bool certain = true;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            // gap between rank k and rank k+1 (only where both exist)
            if (top[r].j[k] != kNone && top[r].j[k + 1] != kNone &&
                !(top[r].d[k] - top[r].d[k + 1] > kBand)) certain = false;
        }
        if (certain) {
            const double sx = x1[i], sy = y1[i], sc = c1[i];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const uint32_t cj = top[r].j[k];
                if (cj == kNone) {
                    id_out[(uint64_t)k*plane_stride + i] = XM_NPOS;
                    d_out[(uint64_t)k*plane_stride + i] = proxy_to_true(dinf(), false);
                } else {
                    const double pr = proxy_sph(sx, sy, sc, x2[cj], y2[cj], c2[cj], sin_model);
                    id_out[(uint64_t)k*plane_stride + i] = (uint64_t)cj;
                    d_out[(uint64_t)k*plane_stride + i] = proxy_to_true(pr, false);
                }
            }
        } else {
            const unsigned long long pos = atomicAdd(refine_count, 1ull);
            refine_list[pos] = i;
            refine_thr[pos] = top[r].j[K - 1] == kNone ? -dinf() : top[r].d[K - 1];
        }
    }
}

// One CTA per flagged row: exact top-K by (proxy, j) among the candidates the filter cannot
// exclude. The list length lives on the device (written by the filter kernel).
template <int K, bool SELF>
__global__ void __launch_bounds__(kThreads)
brute_refine_kernel(int sin_model, const UVec* __restrict__ u1, const UVec* __restrict__ u2, uint32_t n2,
                    const double* __restrict__ x1, const double* __restrict__ y1,
                    const double* __restrict__ c1, const double* __restrict__ x2,
                    const double* __restrict__ y2, const double* __restrict__ c2,
                    uint64_t plane_stride, uint64_t* __restrict__ id_out, double* __restrict__ d_out,
                    const uint32_t* __restrict__ refine_list, const double* __restrict__ refine_thr,
                    const unsigned long long* refine_count, SearchCounters* ctr) {
    __shared__ double s_p[kThreads];
    __shared__ uint32_t s_j[kThreads];
    const unsigned long long nrows = *refine_count;
    for (unsigned long long it = blockIdx.x; it < nrows; it += gridDim.x) {
        const uint32_t i = refine_list[it];
        const double thr = refine_thr[it] - kBand;
        const UVec a = u1[i];
        const double sx = x1[i], sy = y1[i], sc = c1[i];
        double lp[K]; uint32_t lj[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { lp[k] = dinf(); lj[k] = kNone; }
        for (uint32_t j = threadIdx.x; j < n2; j += kThreads) {
            if (SELF && j == i) continue;
            const UVec b = u2[j];
            const double dot = fma(a.ux, b.ux, fma(a.uy, b.uy, a.uz*b.uz));
            if (!(dot >= thr)) continue;
            const double p = proxy_sph(sx, sy, sc, x2[j], y2[j], c2[j], sin_model);
            if (p < lp[K - 1]) {                          // ascending j per thread: strict '<' keeps ties in order
                lp[K - 1] = p; lj[K - 1] = j;
#pragma unroll
                for (int k = K - 2; k >= 0; --k) {
                    if (lp[k] > lp[k + 1]) {
                        const double tp = lp[k]; lp[k] = lp[k + 1]; lp[k + 1] = tp;
                        const uint32_t tj = lj[k]; lj[k] = lj[k + 1]; lj[k + 1] = tj;
                    }
                }
            }
        }
        // K rounds of block-wide argmin over (proxy, j); the winning thread pops its head
        int head = 0;
        for (int k = 0; k < K; ++k) {
            double hp = dinf(); uint32_t hj = kNone;
#pragma unroll
            for (int q = 0; q < K; ++q) if (q == head) { hp = lp[q]; hj = lj[q]; }
            s_p[threadIdx.x] = hp; s_j[threadIdx.x] = hj;
            __syncthreads();
            for (int o = kThreads/2; o > 0; o >>= 1) {
                if ((int)threadIdx.x < o) {
                    const double p2 = s_p[threadIdx.x + o]; const uint32_t j2 = s_j[threadIdx.x + o];
                    const double p1 = s_p[threadIdx.x]; const uint32_t j1 = s_j[threadIdx.x];
                    if (p2 < p1 || (p2 == p1 && j2 < j1)) { s_p[threadIdx.x] = p2; s_j[threadIdx.x] = j2; }
                }
                __syncthreads();
            }
            const double wp = s_p[0]; const uint32_t wj = s_j[0];
            __syncthreads();
            if (wj != kNone && wj == hj) ++head;          // my head won this round
            if (threadIdx.x == 0) {
                id_out[(uint64_t)k*plane_stride + i] = wj == kNone ? XM_NPOS : (uint64_t)wj;
                d_out[(uint64_t)k*plane_stride + i] = proxy_to_true(wp, false);
            }
        }
        __syncthreads();
    }
}

template <int K, int R>
int32_t run_brute_fast(int sin_model, const UVec* u1, uint32_t n1, const UVec* u2, uint32_t n2,
                       const double* x1, const double* y1, const double* c1, const double* x2,
                       const double* y2, const double* c2, bool self, uint64_t stride, uint64_t* id_out,
                       double* d_out, uint32_t* rlist, double* rthr, unsigned long long* rcount,
                       SearchCounters* ctr, cudaStream_t s, Error& err, uint64_t* launches) {
    const dim3 grid((n1 + kThreads*R - 1)/(kThreads*R));
    if (self) {
        brute_filter_kernel<K, R, true><<<grid, kThreads, 0, s>>>(sin_model, u1, n1, u2, n2, x1, y1, c1, x2, y2,
                                                                 c2, stride, id_out, d_out, rlist, rthr, rcount);
        brute_refine_kernel<K, true><<<148*2, kThreads, 0, s>>>(sin_model, u1, u2, n2, x1, y1, c1, x2, y2, c2,
                                                               stride, id_out, d_out, rlist, rthr, rcount, ctr);
    } else {
        brute_filter_kernel<K, R, false><<<grid, kThreads, 0, s>>>(sin_model, u1, n1, u2, n2, x1, y1, c1, x2, y2,
                                                                  c2, stride, id_out, d_out, rlist, rthr, rcount);
        brute_refine_kernel<K, false><<<148*2, kThreads, 0, s>>>(sin_model, u1, u2, n2, x1, y1, c1, x2, y2, c2,
                                                                stride, id_out, d_out, rlist, rthr, rcount, ctr);
    }
    *launches += 2;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

}  // namespace

size_t brute_uvec_bytes(uint64_t n) { return (size_t)n*sizeof(UVec); }

int32_t launch_unit_vectors(const double* x, const double* y, const double* c, uint32_t n, void* out,
                            cudaStream_t s, Error& err, uint64_t* launches) {
    if (n == 0) return XM_OK;
    unit_vectors_kernel<<<(n + 255)/256, 256, 0, s>>>(x, y, c, n, (UVec*)out);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

bool brute_fast_applicable(bool linear, uint32_t nth) { return !linear && nth >= 1 && nth <= 8; }

int32_t launch_brute_fast(int sin_model, const void* u1, uint32_t n1, const void* u2, uint32_t n2,
                          const double* x1, const double* y1, const double* c1, const double* x2,
                          const double* y2, const double* c2, bool self, uint32_t nth,
                          uint64_t plane_stride, uint64_t* id_out, double* d_out, uint32_t* refine_list,
                          double* refine_thr, unsigned long long* refine_count, SearchCounters* ctr,
                          cudaStream_t s, Error& err, uint64_t* launches) {
    if (n1 == 0) return XM_OK;
    const UVec* a = (const UVec*)u1; const UVec* b = (const UVec*)u2;
    switch (nth) {
    case 1: return run_brute_fast<1, 4>(sin_model, a, n1, b, n2, x1, y1, c1, x2, y2, c2, self, plane_stride,
                                        id_out, d_out, refine_list, refine_thr, refine_count, ctr, s, err, launches);
    case 2: return run_brute_fast<2, 2>(sin_model, a, n1, b, n2, x1, y1, c1, x2, y2, c2, self, plane_stride,
                                        id_out, d_out, refine_list, refine_thr, refine_count, ctr, s, err, launches);
    case 3: return run_brute_fast<3, 2>(sin_model, a, n1, b, n2, x1, y1, c1, x2, y2, c2, self, plane_stride,
                                        id_out, d_out, refine_list, refine_thr, refine_count, ctr, s, err, launches);
    case 4: return run_brute_fast<4, 2>(sin_model, a, n1, b, n2, x1, y1, c1, x2, y2, c2, self, plane_stride,
                                        id_out, d_out, refine_list, refine_thr, refine_count, ctr, s, err, launches);
    case 5: return run_brute_fast<5, 1>(sin_model, a, n1, b, n2, x1, y1, c1, x2, y2, c2, self, plane_stride,
                                        id_out, d_out, refine_list, refine_thr, refine_count, ctr, s, err, launches);
    case 6: return run_brute_fast<6, 1>(sin_model, a, n1, b, n2, x1, y1, c1, x2, y2, c2, self, plane_stride,
                                        id_out, d_out, refine_list, refine_thr, refine_count, ctr, s, err, launches);
    case 7: return run_brute_fast<7, 1>(sin_model, a, n1, b, n2, x1, y1, c1, x2, y2, c2, self, plane_stride,
                                        id_out, d_out, refine_list, refine_thr, refine_count, ctr, s, err, launches);
    default: return run_brute_fast<8, 1>(sin_model, a, n1, b, n2, x1, y1, c1, x2, y2, c2, self, plane_stride,
                                         id_out, d_out, refine_list, refine_thr, refine_count, ctr, s, err, launches);
    }
}

}  // namespace xm
