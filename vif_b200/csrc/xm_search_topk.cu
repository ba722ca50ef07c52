// nth-nearest filter kernel for the cell-binned path, 2 <= nth <= 8 (spherical) -- the kernel of
// BASELINE config 3 (self match, nth = 5) and of catalog_merge's nth = 2 calls.
//
// Same idea as ring_nn_filter_kernel (xm_search_fast.cu): rank candidates with the 6-flop metric
// Q = dy^2 + dx^2 cos cos, keep the K+1 smallest by the HIGH WORD of Q (integer compares), decide the
// row only if every adjacent pair of the list is separated by more than the filter's uncertainty,
// evaluate the reference-order proxy of the K winners once; everything else goes to the exact kernel.
//
// What is specific to nth >= 2 is the reference's multiset rule: the ring table
// mirrors one quadrant to the other three, so cells on an axis of ring 1 are visited TWICE and their
// candidates can occupy two ranks. The final list is the K smallest of the multiset in which those
// candidates count twice -- order-independent, so the cells can be visited in any order: here a cell
// on an axis simply inserts each of its candidates twice. Two entries with the same slot are the same
// candidate (an exact tie in the reference as well) and need no separation.
#include "xm_common.cuh"
#include "xm_math_exact.cuh"
#include "xm_filter_common.cuh"

namespace xm {

namespace {

using namespace filt;

constexpr int kThreads = 128;
constexpr int kQueue = 4;                   // ring-1 cells a source may queue; the rest is scanned in place
constexpr int kMaxItems = kQueue*kThreads;
constexpr uint32_t kNoSlot = 0xffffffffu;

// K+1 smallest (high word, slot) pairs, ascending
template <int K>
struct TopK {
    int32_t  h[K + 1];
    uint32_t s[K + 1];
    __device__ __forceinline__ void init(int32_t fill) {
#pragma unroll
        for (int k = 0; k <= K; ++k) { h[k] = fill; s[k] = kNoSlot; }
    }
    __device__ __forceinline__ void insert(int32_t hn, uint32_t slot) {
        if (hn < h[K]) {
            h[K] = hn; s[K] = slot;
#pragma unroll
            for (int k = K - 1; k >= 0; --k) {
                if (h[k] > h[k + 1]) {
                    const int32_t th = h[k]; h[k] = h[k + 1]; h[k + 1] = th;
                    const uint32_t ts = s[k]; s[k] = s[k + 1]; s[k + 1] = ts;
                }
            }
        }
    }
};

template <int K, bool SELF>
__device__ __forceinline__ void scan_cell(const Rec* __restrict__ rec, uint32_t qs, uint32_t qe, double x1,
                                          double y1, double c1, uint32_t self_row, int twice, TopK<K>& L) {
    uint32_t q = qs;
    for (; q + 2 <= qe; q += 2) {
        const Rec r0 = load_rec(rec + q), r1 = load_rec(rec + q + 1);
        const double dx0 = r0.x - x1, dy0 = r0.y - y1, dx1 = r1.x - x1, dy1 = r1.y - y1;
        int32_t h0 = __double2hiint(fma(dx0*dx0*r0.c, c1, dy0*dy0));
        int32_t h1 = __double2hiint(fma(dx1*dx1*r1.c, c1, dy1*dy1));
        if (SELF && r0.idx == self_row) h0 = kInfHi;      // never pair a source with itself (by original row)
        if (SELF && r1.idx == self_row) h1 = kInfHi;
        if (h0 < L.h[K]) { L.insert(h0, q); if (twice) L.insert(h0, q); }
        if (h1 < L.h[K]) { L.insert(h1, q + 1); if (twice) L.insert(h1, q + 1); }
    }
    if (q < qe) {
        const Rec r0 = load_rec(rec + q);
        const double dx0 = r0.x - x1, dy0 = r0.y - y1;
        int32_t h0 = __double2hiint(fma(dx0*dx0*r0.c, c1, dy0*dy0));
        if (SELF && r0.idx == self_row) h0 = kInfHi;
        if (h0 < L.h[K]) { L.insert(h0, q); if (twice) L.insert(h0, q); }
    }
}

template <int K>
struct TopkSmem {
    double   x[kThreads], y[kThreads], c[kThreads];
    uint32_t row[kThreads];                  // original row of each source (self exclusion)
    uint2    range[kMaxItems];
    int32_t  thr[kMaxItems];                 // owner's (K+1)-th high word when the pair was queued
    int32_t  rh[kMaxItems][K + 1];           // partial lists of the queued pairs
    uint32_t rs[kMaxItems][K + 1];
    uint32_t warp_tot[kThreads/32];
    uint8_t  owner[kMaxItems];
    uint8_t  twice[kMaxItems];
};

template <int K, bool SELF>
__global__ void __launch_bounds__(kThreads)
ring_topk_filter_kernel(GridView g, CatView src, CatView cand, uint64_t plane_stride,
                        uint64_t* __restrict__ id_out, double* __restrict__ d_out,
                        uint32_t* __restrict__ refine_list, unsigned long long* refine_count,
                        SearchCounters* ctr) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TopkSmem<K>& sm = *reinterpret_cast<TopkSmem<K>*>(smem_raw);

    const uint32_t p0 = blockIdx.x*blockDim.x;
    const uint32_t p = p0 + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool live = p < src.n;
    const int ndec = (int)g.ndec, nra = (int)g.nra;
    uint32_t evals = 0;

    Rec me;
    me.x = me.y = me.c = 0.0; me.idx = 0; me.key = 0;
    TopK<K> L;
    L.init(kInfHi);
    bool flag = false, done = !live;
    uint32_t mask = 0;
    double cdist = 0.0;

    if (live) {
        me = load_rec(src.rec + p);
        const int x0 = (int)(me.key/g.ndec), y0 = (int)(me.key - (uint32_t)x0*g.ndec);
        const double x1 = me.x, y1 = me.y, c1 = me.c;
        // ---- ring 0: own cell --------------------------------------------------------------------
        {
            const uint32_t qs = cand.cstart[me.key], qe = cand.cstart[me.key + 1];
            scan_cell<K, SELF>(cand.rec, qs, qe, x1, y1, c1, me.idx + g.self_off, 0, L);
            evals += qe - qs;
        }
        const double ex_lo = g.ox + (double)x0*g.wx, ey_lo = g.oy + (double)y0*g.wy;
        const double dxc = (ex_lo + 0.5*g.wx) - x1, dyc = (ey_lo + 0.5*g.wy) - y1;
        cdist = sqrt(fma(dxc*dxc*g.ccen_row[y0], c1, dyc*dyc));
        {   // stop rule after ring 0 on the K-th best (qxmatch.hpp:336-340)
            const int dec = stop_decision(hi_floor(L.h[K-1]), hi_ceil(L.h[K-1]), 0.4*g.cs_rad - 1.1*cdist, g);
            if (dec == 0) done = true; else if (dec == 2) { flag = true; done = true; }
        }
        if (!done) {
            // ---- ring 1: which of the other 24 cells of the 5x5 block can still contribute? -------
            double gl = x1 - ex_lo - g.fzx, gr = (ex_lo + g.wx) - x1 - g.fzx;
            double gd = y1 - ey_lo - g.fzy, gu = (ey_lo + g.wy) - y1 - g.fzy;
            gl = gl > 0 ? gl : 0; gr = gr > 0 ? gr : 0; gd = gd > 0 ? gd : 0; gu = gu > 0 ? gu : 0;
            const double shrink = 1.0 - g.eps_lo - 1e-9;
            double GX[5], GY[5], CM[5];
            GX[2] = 0.0; GX[3] = gr*gr*c1; GX[4] = (gr + g.wx)*(gr + g.wx)*c1;
            GX[1] = gl*gl*c1; GX[0] = (gl + g.wx)*(gl + g.wx)*c1;
            GY[2] = 0.0; GY[3] = gu*gu*shrink; GY[4] = (gu + g.wy)*(gu + g.wy)*shrink;
            GY[1] = gd*gd*shrink; GY[0] = (gd + g.wy)*(gd + g.wy)*shrink;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int yy = y0 + k - 2;
                CM[k] = (yy >= 0 && yy < ndec) ? g.cmin_row[yy]*shrink : -1.0;
            }
            // a cell is useless once its bound reaches the K-th best (candidates there cannot enter
            // ranks 0..K-1); the (K+1)-th entry only serves the ambiguity test
            const double thr = hi_ceil(L.h[K-1])*(1.0 + g.eps_hi)*(1.0 + 1e-9);
#pragma unroll
            for (int e = 0; e < 24; ++e) {
                constexpr int8_t BX[24] = {0, 1, 0, -1, 1, -1, -1, 1,
                                           0, 1, 2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1};
                constexpr int8_t BY[24] = {1, 0, -1, 0, 1, 1, -1, -1,
                                           2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1, 0, 1, 2, 2};
                const int bx = BX[e], by = BY[e];
                const int cx = x0 + bx;
                const double cm = CM[by + 2];
                const double lb = fma(GX[bx + 2], cm, GY[by + 2]);
                if (cx >= 0 && cx < nra && cm >= 0.0 && lb < thr) mask |= 1u << e;
            }
        }
    }

    // ---- queue the surviving non-empty (source, cell) pairs -------------------------------------------
    sm.x[threadIdx.x] = me.x; sm.y[threadIdx.x] = me.y; sm.c[threadIdx.x] = me.c;
    sm.row[threadIdx.x] = me.idx + g.self_off;
    uint2 rng[kQueue];
    uint32_t tw = 0;                         // bit k: queued pair k counts twice
    uint32_t cnt = 0;
    {
        uint32_t m = mask;
        while (m) {
            const int e = __ffs(m) - 1;
            m &= m - 1;
            const int bx = c_bx[e], by = c_by[e];
            const uint32_t ck = (uint32_t)((int)me.key + bx*ndec + by);
            const uint2 r = make_uint2(cand.cstart[ck], cand.cstart[ck + 1]);
            if (r.x == r.y) continue;
            const int twice = (bx == 0 || by == 0) ? 1 : 0;      // axis cells of ring 1 are visited twice
            if (cnt < (uint32_t)kQueue) {
                if (cnt == 0) rng[0] = r; else if (cnt == 1) rng[1] = r; else if (cnt == 2) rng[2] = r; else rng[3] = r;
                tw |= (uint32_t)twice << cnt;
                ++cnt;
            } else {
                scan_cell<K, SELF>(cand.rec, r.x, r.y, me.x, me.y, me.c, me.idx + g.self_off, twice, L);
                evals += r.y - r.x;
            }
        }
    }
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) sm.warp_tot[wid] = inc;
    __syncthreads();
    uint32_t first = inc - cnt, total = 0;
#pragma unroll
    for (int k = 0; k < kThreads/32; ++k) {
        if (k < wid) first += sm.warp_tot[k];
        total += sm.warp_tot[k];
    }
#pragma unroll
    for (int k = 0; k < kQueue; ++k) {
        if ((uint32_t)k < cnt) {
            sm.owner[first + k] = (uint8_t)threadIdx.x; sm.range[first + k] = rng[k];
            sm.twice[first + k] = (uint8_t)((tw >> k) & 1u); sm.thr[first + k] = L.h[K];
        }
    }
    __syncthreads();

    // ---- one queued pair per thread: partial list of the K+1 smallest below the owner's threshold ----
    for (uint32_t it = threadIdx.x; it < total; it += kThreads) {
        const int owner = sm.owner[it];
        const uint2 r = sm.range[it];
        TopK<K> A;
        A.init(sm.thr[it]);
        scan_cell<K, SELF>(cand.rec, r.x, r.y, sm.x[owner], sm.y[owner], sm.c[owner], sm.row[owner], 0, A);
        evals += r.y - r.x;
#pragma unroll
        for (int k = 0; k <= K; ++k) { sm.rh[it][k] = A.h[k]; sm.rs[it][k] = A.s[k]; }
    }
    __syncthreads();
    for (uint32_t k = 0; k < cnt; ++k) {
        const uint32_t it = first + k;
        const int twice = sm.twice[it];
#pragma unroll
        for (int e = 0; e <= K; ++e) {
            const uint32_t sl = sm.rs[it][e];
            if (sl != kNoSlot) { L.insert(sm.rh[it][e], sl); if (twice) L.insert(sm.rh[it][e], sl); }
        }
    }

    // ---- decide -------------------------------------------------------------------------------------------
    if (live) {
        if (!done) {
            const int dec = stop_decision(hi_floor(L.h[K-1]), hi_ceil(L.h[K-1]), 2.4*g.cs_rad - 1.1*cdist, g);
            if (dec != 0) flag = true;           // ring 2 and beyond: the exact kernel's job
        }
        // every rank must exist, and neighbours in the list must be separated by more than the
        // filter's uncertainty unless they are the same candidate (counted twice)
        if (L.s[K-1] == kNoSlot) flag = true;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (L.s[k + 1] != kNoSlot && L.s[k] != L.s[k + 1] &&
                !(hi_ceil(L.h[k])*(1.0 + g.eps_hi) < hi_floor(L.h[k + 1])*(1.0 - g.eps_lo))) flag = true;
        }
        if (flag) {
            const unsigned long long pos = atomicAdd(refine_count, 1ull);
            refine_list[pos] = p;
        } else {
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const Rec w = load_rec(cand.rec + L.s[k]);
                const double pr = proxy_sph_small(me.x, me.y, me.c, w.x, w.y, w.c, g.sin_model);
                id_out[(uint64_t)k*plane_stride + me.idx] = (uint64_t)w.idx;
                d_out[(uint64_t)k*plane_stride + me.idx] = proxy_to_true_small(pr);
            }
        }
    }
    unsigned long long v = evals;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v) atomicAdd(&ctr->evals_fwd, v);
}

template <int K>
int32_t run_topk(const GridView& g, const CatView& src, const CatView& cand, bool self, uint64_t stride,
                 uint64_t* id_out, double* d_out, uint32_t* refine_list, unsigned long long* refine_count,
                 SearchCounters* ctr, cudaStream_t s, Error& err) {
    const uint32_t blocks = (src.n + kThreads - 1)/kThreads;
    const size_t smem = sizeof(TopkSmem<K>);
    if (self) {
        XM_CUDA_TRY(cudaFuncSetAttribute(ring_topk_filter_kernel<K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ring_topk_filter_kernel<K, true><<<blocks, kThreads, smem, s>>>(g, src, cand, stride, id_out, d_out, refine_list,
                                                                       refine_count, ctr);
    } else {
        XM_CUDA_TRY(cudaFuncSetAttribute(ring_topk_filter_kernel<K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ring_topk_filter_kernel<K, false><<<blocks, kThreads, smem, s>>>(g, src, cand, stride, id_out, d_out, refine_list,
                                                                        refine_count, ctr);
    }
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

}  // namespace

bool topk_filter_applicable(uint32_t nth_eff) { return nth_eff >= 2 && nth_eff <= 8; }

// Forward search only (the reverse match is always nth = 1). Rows it cannot decide are appended to
// refine_list / *refine_count for ring_search_exact_kernel (which needs cells in row order).
int32_t launch_ring_topk_filter(const GridView& g, const CatView& src, const CatView& cand, bool self,
                                uint32_t nth_eff, uint64_t plane_stride, uint64_t* id_out, double* d_out,
                                uint32_t* refine_list, unsigned long long* refine_count, SearchCounters* ctr,
                                cudaStream_t s, Error& err, uint64_t* launches) {
    if (src.n == 0) return XM_OK;
    *launches += 1;
    switch (nth_eff) {
    case 2: return run_topk<2>(g, src, cand, self, plane_stride, id_out, d_out, refine_list, refine_count, ctr, s, err);
    case 3: return run_topk<3>(g, src, cand, self, plane_stride, id_out, d_out, refine_list, refine_count, ctr, s, err);
    case 4: return run_topk<4>(g, src, cand, self, plane_stride, id_out, d_out, refine_list, refine_count, ctr, s, err);
    case 5: return run_topk<5>(g, src, cand, self, plane_stride, id_out, d_out, refine_list, refine_count, ctr, s, err);
    case 6: return run_topk<6>(g, src, cand, self, plane_stride, id_out, d_out, refine_list, refine_count, ctr, s, err);
    case 7: return run_topk<7>(g, src, cand, self, plane_stride, id_out, d_out, refine_list, refine_count, ctr, s, err);
    default: return run_topk<8>(g, src, cand, self, plane_stride, id_out, d_out, refine_list, refine_count, ctr, s, err);
    }
}

}  // namespace xm
