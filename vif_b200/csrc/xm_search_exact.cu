// Exact searches: every distance is evaluated with the reference's expression order
// (xm_math_exact.cuh). These kernels define the results; the filter kernels of xm_search_fast.cu
// only decide which rows may skip them.
//
//   ring_search_exact  = work1 / work2 of the reference (qxmatch.hpp:293-380): rings outward from
//                        the source's own cell in the depth_cache order, ascending row order inside
//                        a cell (the index is a stable sort), strict-'<' insertion, the reference's
//                        stop rule. One thread per source, sources in cell order so that a warp's
//                        candidate loads hit the same few cells.
//   brute_exact        = the brute_force path (qxmatch.hpp:491-526), one thread per source, all
//                        candidates in ascending order through shared-memory tiles.
//
// Culling: a cell is skipped when a rigorous lower bound of the proxy over the cell is >= the
// current nth-best. The reference would have evaluated those candidates and rejected every one of
// them with the same strict '<', so the lists are unchanged.
#include "xm_common.cuh"
#include "xm_math_exact.cuh"
#include "xm_ring.cuh"

namespace xm {

namespace {

constexpr int kRingThreads = 128;
constexpr int kBruteThreads = 256;
constexpr uint32_t kNoCand = 0xffffffffu;

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000ll); }

// ---- nth-best lists -------------------------------------------------------------------------
// insertion rule of qxmatch.hpp:323-332: replace the last slot if strictly smaller, then bubble
// towards the front while the earlier entry is strictly larger (equal earlier entries stay first).
template <int K>
struct RegList {
    static constexpr int kK = K;
    double   d[K];
    uint32_t q[K];
    double   rej;      // smallest proxy seen that is NOT in the list (ambiguity diagnostic)
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int k = 0; k < K; ++k) { d[k] = dinf(); q[k] = kNoCand; }
        rej = dinf();
    }
    __device__ __forceinline__ double kth() const { return d[K-1]; }
    __device__ __forceinline__ void insert(double sd, uint32_t cq) {
        if (sd < d[K-1]) {
            rej = fmin(rej, d[K-1]);
            d[K-1] = sd; q[K-1] = cq;
#pragma unroll
            for (int k = K-2; k >= 0; --k) {
                if (d[k] > d[k+1]) {
                    const double td = d[k]; d[k] = d[k+1]; d[k+1] = td;
                    const uint32_t tq = q[k]; q[k] = q[k+1]; q[k+1] = tq;
                }
            }
        } else {
            rej = fmin(rej, sd);
        }
    }
    __device__ __forceinline__ double get_d(int k) const { return d[k]; }
    __device__ __forceinline__ uint32_t get_q(int k) const { return q[k]; }
    __device__ __forceinline__ void set0(double sd, uint32_t cq) { d[0] = sd; q[0] = cq; }
};

// nth > 8: the list lives in the output planes themselves (proxy in d, candidate slot in id)
struct MemList {
    static constexpr int kK = 0;
    double*   d;
    uint64_t* q;
    uint64_t  stride;
    uint32_t  k_eff;
    double    rej;
    __device__ __forceinline__ void init() {
        for (uint32_t k = 0; k < k_eff; ++k) { d[k*stride] = dinf(); q[k*stride] = kNoCand; }
        rej = dinf();
    }
    __device__ __forceinline__ double kth() const { return d[(uint64_t)(k_eff-1)*stride]; }
    __device__ __forceinline__ void insert(double sd, uint32_t cq) {
        const uint64_t last = (uint64_t)(k_eff-1)*stride;
        if (sd < d[last]) {
            rej = fmin(rej, d[last]);
            d[last] = sd; q[last] = cq;
            for (uint32_t k = k_eff-1; k-- > 0; ) {
                double* a = d + (uint64_t)k*stride; double* b = a + stride;
                if (!(*a > *b)) break;
                const double td = *a; *a = *b; *b = td;
                uint64_t* qa = q + (uint64_t)k*stride; uint64_t* qb = qa + stride;
                const uint64_t tq = *qa; *qa = *qb; *qb = tq;
            }
        } else {
            rej = fmin(rej, sd);
        }
    }
    __device__ __forceinline__ double get_d(int k) const { return d[(uint64_t)k*stride]; }
    __device__ __forceinline__ uint32_t get_q(int k) const { return (uint32_t)q[(uint64_t)k*stride]; }
    __device__ __forceinline__ void set0(double sd, uint32_t cq) { d[0] = sd; q[0] = cq; }
};

// 0 < |a-b| <= 8 ulp: an ordering that a libm differing in the last bits could flip
__device__ __forceinline__ bool near_tie(double a, double b) {
    if (!(a < dinf()) || !(b < dinf()) || a == b) return false;
    return fabs(a - b) <= 1.8e-15*fmax(fabs(a), fabs(b));
}

template <class List>
__device__ __forceinline__ bool list_ambiguous(const List& L, uint32_t k_eff) {
    bool amb = false;
    for (uint32_t k = 0; k + 1 < k_eff; ++k) amb |= near_tie(L.get_d(k), L.get_d(k+1));
    amb |= near_tie(L.get_d(k_eff-1), L.rej);
    return amb;
}

// all 32 lanes of every warp must reach this (no early returns in the kernels)
__device__ __forceinline__ void warp_add(unsigned long long* dst, unsigned long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(dst, v);
}

// ---- rigorous lower bound of the proxy over one cell -----------------------------------------
__device__ __forceinline__ double sin2_lower(double h) {       // h >= 0
    if (h < 0.02) { const double hh = h*h; return hh*(1.0 - hh*(1.0/3.0)); }
    if (h <= 3.2) { const double s = sin(h); return s*s; }
    return 0.0;
}

template <bool LINEAR>
__device__ __forceinline__ double cell_lower_bound(const GridView& g, double x1, double y1, double c1,
                                                   int64_t cx, int64_t cy) {
    const double ex_lo = g.ox + (double)cx*g.wx, ex_hi = ex_lo + g.wx;
    const double ey_lo = g.oy + (double)cy*g.wy, ey_hi = ey_lo + g.wy;
    double gx = fmax(ex_lo - x1, x1 - ex_hi) - g.fzx; gx = gx > 0.0 ? gx : 0.0;
    double gy = fmax(ey_lo - y1, y1 - ey_hi) - g.fzy; gy = gy > 0.0 ? gy : 0.0;
    if (LINEAR) return (gx*gx + gy*gy)*(1.0 - 1e-12);
    const double cm = g.cmin_row[cy];
    if (cm < 0.0 || c1 < 0.0) return -dinf();                   // declinations beyond the poles
    const double far_x = fmax(fabs(ex_lo - x1), fabs(ex_hi - x1)) + g.fzx;
    const double a = sin2_lower(0.5*gy);
    const double b = fmin(sin2_lower(0.5*gx), sin2_lower(0.5*far_x));   // sin^2 is unimodal on [0,pi]
    return (a + b*c1*cm)*(1.0 - 1e-9);
}

// ---- ring search ---------------------------------------------------------------------------------
// UNORD (nth = 1 only): the cells of `cand` are in arbitrary order (bucket index without the
// ordering pass). The reference visits a cell in ascending row order and keeps the first of
// equal proxies; the same outcome is obtained by letting an equal proxy from the SAME cell visit
// win when its row is smaller.
template <class List, bool LINEAR, bool UNORD>
__device__ __forceinline__ void ring_search_row(const GridView& g, const CatView& src,
                                                const CatView& cand, int self_mode, uint32_t p,
                                                List& L, unsigned long long& evals) {
    const Rec me = src.rec[p];
    const double x1 = me.x, y1 = me.y, c1 = me.c;
    const uint32_t key = me.key;
    const int64_t x0 = key/g.ndec, y0 = key - (uint32_t)x0*g.ndec;

    // distance of the source from its cell centre (qxmatch.hpp:298-299)
    const double ccx = __dadd_rn(g.rra0,  __dmul_rn(__dadd_rn((double)x0, 0.5), g.dra));
    const double ccy = __dadd_rn(g.rdec0, __dmul_rn(__dadd_rn((double)y0, 0.5), g.ddec));
    double cell_dist;
    if (LINEAR) {
        const double dx = __dsub_rn(ccx, x1), dy = __dsub_rn(ccy, y1);
        cell_dist = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    } else {
        cell_dist = angdist_pre(x1, y1, c1, ccx, ccy, g.sin_model);
    }

    const bool dedupe = (List::kK == 1);
    uint32_t d = 0;
    double reached;
    do {
        for_each_ring_entry(d, g.nra, g.ndec, dedupe, [&](int bx, int by) {
            const int64_t cx = x0 + bx, cy = y0 + by;
            if (cx < 0 || cx >= (int64_t)g.nra || cy < 0 || cy >= (int64_t)g.ndec) return;
            const uint64_t ck = (uint64_t)cx*g.ndec + (uint64_t)cy;
            const uint2 r = make_uint2(cand.cstart[ck], cand.cstart[ck + 1]);
            if (r.x == r.y) return;
            if (d > 0 && cell_lower_bound<LINEAR>(g, x1, y1, c1, cx, cy) >= L.kth()) return;
            bool best_here = false;            // UNORD: the current best came from this cell visit
            uint32_t best_row = 0;
            for (uint32_t q = r.x; q < r.y; ++q) {
                // self match (qxmatch.hpp:314): skip the source itself -- same slot when both
                // catalogs share one index (1), same original row otherwise (2)
                if (self_mode == 1 && q == p) continue;
                const Rec cd = cand.rec[q];
                if (self_mode == 2 && cd.idx == me.idx + g.self_off) continue;
                const double sd = LINEAR ? proxy_lin(x1, y1, cd.x, cd.y)
                                         : proxy_sph(x1, y1, c1, cd.x, cd.y, cd.c, g.sin_model);
                ++evals;
                if (UNORD) {
                    const double cur = L.get_d(0);
                    if (sd < cur || (sd == cur && best_here && cd.idx < best_row)) {
                        L.insert(sd < cur ? sd : -1.0, q);      // forced replacement of slot 0 ...
                        L.set0(sd, q);                          // ... with the true proxy
                        best_here = true; best_row = cd.idx;
                    } else {
                        L.insert(sd, q);                        // rejected: only updates the diagnostics
                    }
                } else {
                    L.insert(sd, q);
                }
            }
        });
        // qxmatch.hpp:336-337
        const double max_dist = d == 0 ? __ddiv_rn(g.cell_size, 2.0)
                                       : __dmul_rn(g.cell_size, __dadd_rn((double)(d + 1), 0.5));
        const double t = __dsub_rn(__dsub_rn(max_dist, __dmul_rn(1.1, cell_dist)),
                                   __dmul_rn(0.1, g.cell_size));
        reached = true_to_proxy(t > 0.0 ? t : 0.0, LINEAR, g.sin_model);
        ++d;
    } while (L.kth() > reached && d < g.max_depth);
}

template <int K, bool LINEAR>
__global__ void __launch_bounds__(kRingThreads)
ring_search_exact_kernel(GridView g, CatView src, CatView cand, int self_mode, uint32_t nth_eff,
                         uint64_t plane_stride, uint64_t* __restrict__ id_out,
                         double* __restrict__ d_out, const uint32_t* __restrict__ rows,
                         uint32_t nrows, const unsigned long long* __restrict__ nrows_dev,
                         int reverse, int unordered, SearchCounters* ctr) {
    // rows == nullptr: thread t handles sorted slot t. Otherwise the kernel walks the row list with
    // a grid-stride loop; its length may live on the device (filter kernel's refine count).
    if (nrows_dev) { const unsigned long long c = *nrows_dev; nrows = c < nrows ? (uint32_t)c : nrows; }
    unsigned long long evals = 0, amb = 0;
    const uint32_t stride = gridDim.x*blockDim.x;
    for (uint32_t t = blockIdx.x*blockDim.x + threadIdx.x; t < nrows; t += stride) {
        const uint32_t p = rows ? rows[t] : t;
        const uint32_t i = src.rec[p].idx;
        if (K > 0) {
            RegList<(K > 0 ? K : 1)> L;
            L.init();
            if (K == 1 && unordered) ring_search_row<RegList<(K > 0 ? K : 1)>, LINEAR, (K == 1)>(g, src, cand, self_mode, p, L, evals);
            else ring_search_row<RegList<(K > 0 ? K : 1)>, LINEAR, false>(g, src, cand, self_mode, p, L, evals);
#pragma unroll
            for (int k = 0; k < (K > 0 ? K : 1); ++k) {
                const uint32_t q = L.q[k];
                id_out[(uint64_t)k*plane_stride + i] = q == kNoCand ? XM_NPOS : (uint64_t)cand.rec[q].idx;
                d_out[(uint64_t)k*plane_stride + i] = proxy_to_true(L.d[k], LINEAR);
            }
            amb += list_ambiguous(L, (uint32_t)(K > 0 ? K : 1)) ? 1 : 0;
        } else {
            MemList L;
            L.d = d_out + i; L.q = id_out + i; L.stride = plane_stride; L.k_eff = nth_eff;
            L.init();
            ring_search_row<MemList, LINEAR, false>(g, src, cand, self_mode, p, L, evals);
            amb += list_ambiguous(L, nth_eff) ? 1 : 0;
            for (uint32_t k = 0; k < nth_eff; ++k) {
                const uint32_t q = L.get_q(k);
                const double pd = L.get_d(k);
                id_out[(uint64_t)k*plane_stride + i] = q == kNoCand ? XM_NPOS : (uint64_t)cand.rec[q].idx;
                d_out[(uint64_t)k*plane_stride + i] = proxy_to_true(pd, LINEAR);
            }
        }
    }
    __syncwarp();
    warp_add(reverse ? &ctr->evals_rev : &ctr->evals_fwd, evals);
    warp_add(&ctr->ambiguous, amb);
}

// ---- list mode, nth = 1: one WARP per flagged row ------------------------------------------------
// The filter kernels hand over a few thousand rows per call. One thread per row (kernel above) is a
// serial chain of ~60 dependent global loads per row (cell offsets, then candidates): ~130 us whatever
// the list length. Here a warp owns a row: lane l resolves ring-1 entry l (range + lower bound) in
// parallel, surviving cells are visited in the reference's order and each cell's candidates are
// evaluated 32 at a time. Semantics are those of ring_search_row with a one-slot list:
// strictly smaller proxies replace the best, so among equal proxies the first visited cell wins and,
// inside a cell, the first row (ascending slot when cells are ordered, ascending original row when
// they are not). Rows that must go beyond ring 1 fall back to ring_search_row on lane 0.
struct WarpBest {
    double   sd;       // best proxy
    uint32_t q;        // its slot
    double   second;   // smallest other proxy seen (ambiguity diagnostic, = RegList::rej)
};

template <bool UNORD>
__device__ __forceinline__ void warp_scan_cell(const GridView& g, const CatView& cand, const Rec& me, int self_mode,
                                               uint32_t p, uint32_t qs, uint32_t qe, WarpBest& B,
                                               unsigned long long& evals) {
    const int lane = threadIdx.x & 31;
    double m1 = dinf(), m2 = dinf();
    uint32_t mq = kNoCand, mord = 0xffffffffu;
    for (uint32_t q = qs + lane; q < qe; q += 32) {
        if (self_mode == 1 && q == p) continue;
        const Rec cd = cand.rec[q];
        if (self_mode == 2 && cd.idx == me.idx + g.self_off) continue;
        const double sd = proxy_sph(me.x, me.y, me.c, cd.x, cd.y, cd.c, g.sin_model);
        ++evals;
        const uint32_t ord = UNORD ? cd.idx : q;
        if (sd < m1 || (sd == m1 && ord < mord)) { m2 = m1; m1 = sd; mq = q; mord = ord; }
        else m2 = fmin(m2, sd);
    }
    // lexicographic (proxy, order) minimum over the warp + the second smallest proxy
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double a1 = __shfl_xor_sync(0xffffffffu, m1, o), a2 = __shfl_xor_sync(0xffffffffu, m2, o);
        const uint32_t aq = __shfl_xor_sync(0xffffffffu, mq, o), ao = __shfl_xor_sync(0xffffffffu, mord, o);
        const bool take = a1 < m1 || (a1 == m1 && ao < mord);
        const double lose = take ? m1 : a1;
        m2 = fmin(fmin(m2, a2), lose);
        if (take) { m1 = a1; mq = aq; mord = ao; }
    }
    if (mq == kNoCand) return;
    if (m1 < B.sd) { B.second = fmin(fmin(B.second, B.sd), m2); B.sd = m1; B.q = mq; }
    else B.second = fmin(B.second, m1);          // m2 >= m1: m1 is the smallest rejected value of this cell
}

template <bool UNORD>
__global__ void __launch_bounds__(kRingThreads)
ring_list_nn_kernel(GridView g, CatView src, CatView cand, int self_mode, uint64_t* __restrict__ id_out,
                    double* __restrict__ d_out, const uint32_t* __restrict__ rows, uint32_t nrows,
                    const unsigned long long* __restrict__ nrows_dev, int reverse, SearchCounters* ctr) {
    if (nrows_dev) { const unsigned long long c = *nrows_dev; nrows = c < nrows ? (uint32_t)c : nrows; }
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x*blockDim.x) >> 5;
    unsigned long long evals = 0, amb = 0;
    // ring-1 entry of this lane in the reference's order (axis cells once: a repeated visit cannot change a
    // nearest-neighbour search with strict '<')
    int mybx = 0, myby = 0, nent = 0;
    if (g.max_depth > 1)
        for_each_ring_entry(1, g.nra, g.ndec, true, [&](int bx, int by) { if (nent == lane) { mybx = bx; myby = by; } ++nent; });

    for (uint32_t t = warp; t < nrows; t += nwarps) {
        const uint32_t p = rows[t];
        const Rec me = src.rec[p];
        const int64_t x0 = me.key/g.ndec, y0 = me.key - (uint32_t)x0*g.ndec;
        const double ccx = __dadd_rn(g.rra0,  __dmul_rn(__dadd_rn((double)x0, 0.5), g.dra));
        const double ccy = __dadd_rn(g.rdec0, __dmul_rn(__dadd_rn((double)y0, 0.5), g.ddec));
        const double cell_dist = angdist_pre(me.x, me.y, me.c, ccx, ccy, g.sin_model);
        WarpBest B;
        B.sd = dinf(); B.q = kNoCand; B.second = dinf();
        // ring 0
        warp_scan_cell<UNORD>(g, cand, me, self_mode, p, cand.cstart[me.key], cand.cstart[me.key + 1], B, evals);
        auto reached_after = [&](uint32_t d) {
            const double max_dist = d == 0 ? __ddiv_rn(g.cell_size, 2.0) : __dmul_rn(g.cell_size, __dadd_rn((double)(d + 1), 0.5));
            const double tt = __dsub_rn(__dsub_rn(max_dist, __dmul_rn(1.1, cell_dist)), __dmul_rn(0.1, g.cell_size));
            return true_to_proxy(tt > 0.0 ? tt : 0.0, false, g.sin_model);
        };
        bool deeper = false;
        if (B.sd > reached_after(0) && 1 < g.max_depth) {
            // ring 1: every lane resolves its entry, then the surviving cells are visited in order
            uint32_t qs = 0, qe = 0;
            double lb = dinf();
            if (lane < nent) {
                const int64_t cx = x0 + mybx, cy = y0 + myby;
                if (cx >= 0 && cx < (int64_t)g.nra && cy >= 0 && cy < (int64_t)g.ndec) {
                    const uint64_t ck = (uint64_t)cx*g.ndec + (uint64_t)cy;
                    qs = cand.cstart[ck]; qe = cand.cstart[ck + 1];
                    if (qs < qe) lb = cell_lower_bound<false>(g, me.x, me.y, me.c, cx, cy);
                }
            }
            unsigned todo = __ballot_sync(0xffffffffu, qs < qe);
            while (todo) {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1;
                const double clb = __shfl_sync(0xffffffffu, lb, l);
                const uint32_t cqs = __shfl_sync(0xffffffffu, qs, l), cqe = __shfl_sync(0xffffffffu, qe, l);
                if (clb >= B.sd) continue;                          // culled exactly as in ring_search_row
                warp_scan_cell<UNORD>(g, cand, me, self_mode, p, cqs, cqe, B, evals);
            }
            deeper = B.sd > reached_after(1) && 2 < g.max_depth;
        }
        if (deeper) {
            // beyond ring 1 (sparse neighbourhoods): the general walk, from scratch, on one lane
            if (lane == 0) {
                RegList<1> L;
                L.init();
                unsigned long long ev = 0;
                ring_search_row<RegList<1>, false, UNORD>(g, src, cand, self_mode, p, L, ev);
                evals += ev;
                B.sd = L.d[0]; B.q = L.q[0]; B.second = L.rej;
            }
        }
        if (lane == 0) {
            id_out[me.idx] = B.q == kNoCand ? XM_NPOS : (uint64_t)cand.rec[B.q].idx;
            d_out[me.idx] = proxy_to_true(B.sd, false);
            amb += near_tie(B.sd, B.second) ? 1 : 0;
        }
    }
    __syncwarp();
    warp_add(reverse ? &ctr->evals_rev : &ctr->evals_fwd, evals);
    warp_add(&ctr->ambiguous, amb);
}

// ---- brute force -------------------------------------------------------------------------------
template <int K, bool LINEAR>
__global__ void __launch_bounds__(kBruteThreads)
brute_exact_kernel(int sin_model, const double* __restrict__ x1, const double* __restrict__ y1,
                   const double* __restrict__ c1, uint32_t n1, const double* __restrict__ x2,
                   const double* __restrict__ y2, const double* __restrict__ c2, uint32_t n2,
                   int self_mode, uint32_t nth_eff, uint64_t plane_stride,
                   uint64_t* __restrict__ id_out, double* __restrict__ d_out,
                   const uint32_t* __restrict__ rows, uint32_t nrows, int reverse,
                   SearchCounters* ctr) {
    __shared__ double sx[kBruteThreads], sy[kBruteThreads], sc[kBruteThreads];
    const uint32_t t = blockIdx.x*blockDim.x + threadIdx.x;
    const bool active = t < nrows;
    const uint32_t i = active ? (rows ? rows[t] : t) : 0u;
    const double ax = active ? x1[i] : 0.0, ay = active ? y1[i] : 0.0, ac = active ? c1[i] : 0.0;
    unsigned long long evals = 0, amb = 0;

    RegList<(K > 0 ? K : 1)> R;
    MemList M;
    if (K > 0) R.init();
    else if (active) {
        M.d = d_out + i; M.q = id_out + i; M.stride = plane_stride; M.k_eff = nth_eff; M.init();
    }

    for (uint32_t base = 0; base < n2; base += kBruteThreads) {
        const uint32_t j = base + threadIdx.x;
        if (j < n2) { sx[threadIdx.x] = x2[j]; sy[threadIdx.x] = y2[j]; sc[threadIdx.x] = c2[j]; }
        __syncthreads();
        if (active) {
            const uint32_t cnt = min((uint32_t)kBruteThreads, n2 - base);
            for (uint32_t jj = 0; jj < cnt; ++jj) {
                const uint32_t cj = base + jj;
                if (self_mode && cj == i) continue;
                const double sd = LINEAR ? proxy_lin(ax, ay, sx[jj], sy[jj])
                                         : proxy_sph(ax, ay, ac, sx[jj], sy[jj], sc[jj], sin_model);
                if (K > 0) R.insert(sd, cj); else M.insert(sd, cj);
            }
            evals += cnt;
        }
        __syncthreads();
    }
    if (active) {
        if (K > 0) {
#pragma unroll
            for (int k = 0; k < (K > 0 ? K : 1); ++k) {
                id_out[(uint64_t)k*plane_stride + i] = R.q[k] == kNoCand ? XM_NPOS : (uint64_t)R.q[k];
                d_out[(uint64_t)k*plane_stride + i] = proxy_to_true(R.d[k], LINEAR);
            }
            amb = list_ambiguous(R, (uint32_t)(K > 0 ? K : 1)) ? 1 : 0;
        } else {
            amb = list_ambiguous(M, nth_eff) ? 1 : 0;
            for (uint32_t k = 0; k < nth_eff; ++k) {
                const uint32_t q = M.get_q(k);
                const double pd = M.get_d(k);
                id_out[(uint64_t)k*plane_stride + i] = q == kNoCand ? XM_NPOS : (uint64_t)q;
                d_out[(uint64_t)k*plane_stride + i] = proxy_to_true(pd, LINEAR);
            }
        }
    }
    __syncwarp();
    warp_add(reverse ? &ctr->evals_rev : &ctr->evals_fwd, evals);
    warp_add(&ctr->ambiguous, amb);
}

__global__ void prep_unsorted_kernel(const double* __restrict__ ra, const double* __restrict__ dec,
                                     uint32_t n, int linear, int model, double* __restrict__ x,
                                     double* __restrict__ y, double* __restrict__ c) {
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double r = ra[i], d = dec[i];
    if (linear) { x[i] = r; y[i] = d; c[i] = 1.0; }
    else { const double yr = __dmul_rn(d, XM_D2R); x[i] = __dmul_rn(r, XM_D2R); y[i] = yr; c[i] = cos_ref(yr, model); }
}

__global__ void fill_outputs_kernel(uint64_t* __restrict__ id, double* __restrict__ d, uint64_t n,
                                    double dval) {
    const uint64_t stride = (uint64_t)gridDim.x*blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x*blockDim.x + threadIdx.x; i < n; i += stride) {
        id[i] = XM_NPOS; d[i] = dval;
    }
}

// per cell row: lower bound of cos(dec) over the row (-1 where the row is not inside
// [-pi/2, pi/2]) and cos(dec) at the row centre
__global__ void row_tables_kernel(GridView g, double* __restrict__ cmin, double* __restrict__ ccen) {
    const uint32_t cy = blockIdx.x*blockDim.x + threadIdx.x;
    if (cy >= g.ndec) return;
    const double lo = g.oy + (double)cy*g.wy - g.fzy, hi = lo + g.wy + 2.0*g.fzy;
    const double half_pi = 1.5707963267948966;
    double v = -1.0;
    if (lo >= -half_pi && hi <= half_pi) {
        v = fmin(cos(lo), cos(hi))*(1.0 - 1e-12) - 1e-15;
        if (v < 0.0) v = 0.0;
    }
    cmin[cy] = v;
    // centre declination as the reference computes it (qxmatch.hpp:299), then to radians
    const double cdeg = __dadd_rn(g.rdec0, __dmul_rn(__dadd_rn((double)cy, 0.5), g.ddec));
    ccen[cy] = cos(__dmul_rn(XM_D2R, cdeg));
}

template <bool LINEAR>
void dispatch_ring(int K, dim3 grid, cudaStream_t s, const GridView& g, const CatView& src,
                   const CatView& cand, int self_mode, uint32_t nth_eff, uint64_t stride,
                   uint64_t* id_out, double* d_out, const uint32_t* rows, uint32_t nrows,
                   const unsigned long long* nrows_dev, int reverse, int unordered, SearchCounters* ctr) {
#define XM_CASE(KK) case KK: ring_search_exact_kernel<KK, LINEAR><<<grid, kRingThreads, 0, s>>>( \
        g, src, cand, self_mode, nth_eff, stride, id_out, d_out, rows, nrows, nrows_dev, reverse, unordered, ctr); break;
    switch (K) { XM_CASE(1) XM_CASE(2) XM_CASE(3) XM_CASE(4) XM_CASE(5) XM_CASE(6) XM_CASE(7) XM_CASE(8)
                 default: ring_search_exact_kernel<0, LINEAR><<<grid, kRingThreads, 0, s>>>(
                     g, src, cand, self_mode, nth_eff, stride, id_out, d_out, rows, nrows, nrows_dev, reverse, unordered, ctr); }
#undef XM_CASE
}

template <bool LINEAR>
void dispatch_brute(int K, dim3 grid, cudaStream_t s, int sin_model, const double* x1,
                    const double* y1, const double* c1, uint32_t n1, const double* x2,
                    const double* y2, const double* c2, uint32_t n2, int self_mode, uint32_t nth_eff,
                    uint64_t stride, uint64_t* id_out, double* d_out, const uint32_t* rows,
                    uint32_t nrows, int reverse, SearchCounters* ctr) {
#define XM_CASE(KK) case KK: brute_exact_kernel<KK, LINEAR><<<grid, kBruteThreads, 0, s>>>( \
        sin_model, x1, y1, c1, n1, x2, y2, c2, n2, self_mode, nth_eff, stride, id_out, d_out, rows, \
        nrows, reverse, ctr); break;
    switch (K) { XM_CASE(1) XM_CASE(2) XM_CASE(3) XM_CASE(4) XM_CASE(5) XM_CASE(6) XM_CASE(7) XM_CASE(8)
                 default: brute_exact_kernel<0, LINEAR><<<grid, kBruteThreads, 0, s>>>(
                     sin_model, x1, y1, c1, n1, x2, y2, c2, n2, self_mode, nth_eff, stride, id_out,
                     d_out, rows, nrows, reverse, ctr); }
#undef XM_CASE
}

}  // namespace

int32_t launch_ring_search_exact(const GridView& g, const CatView& src, const CatView& cand,
                                 int self, uint32_t nth_eff, uint64_t plane_stride,
                                 uint64_t* id_out, double* d_out, const uint32_t* rows,
                                 uint32_t nrows, const unsigned long long* nrows_dev, bool reverse,
                                 bool cells_unordered, SearchCounters* ctr, cudaStream_t s,
                                 Error& err, uint64_t* launches) {
    if (nrows == 0 || nth_eff == 0) return XM_OK;
    const int unordered = (cells_unordered && nth_eff == 1) ? 1 : 0;
    if (rows && nth_eff == 1 && !g.linear) {
        // list of flagged rows of a nearest-neighbour search: one warp per row
        const uint64_t want = ((uint64_t)nrows*32 + kRingThreads - 1)/kRingThreads;
        const uint32_t nb = (uint32_t)(want < 148ull*16 ? want : 148ull*16);
        if (unordered) ring_list_nn_kernel<true><<<nb, kRingThreads, 0, s>>>(g, src, cand, self, id_out, d_out, rows, nrows,
                                                                              nrows_dev, reverse ? 1 : 0, ctr);
        else ring_list_nn_kernel<false><<<nb, kRingThreads, 0, s>>>(g, src, cand, self, id_out, d_out, rows, nrows,
                                                                    nrows_dev, reverse ? 1 : 0, ctr);
        *launches += 1;
        XM_CUDA_TRY(cudaGetLastError());
        return XM_OK;
    }
    uint32_t blocks = (nrows + kRingThreads - 1)/kRingThreads;
    if (nrows_dev && blocks > 148*4) blocks = 148*4;     // list of unknown (small) length: persistent grid
    const dim3 grid(blocks);
    const int K = nth_eff <= 8 ? (int)nth_eff : 0;
    if (g.linear) dispatch_ring<true>(K, grid, s, g, src, cand, self, nth_eff, plane_stride, id_out,
                                      d_out, rows, nrows, nrows_dev, reverse, unordered, ctr);
    else dispatch_ring<false>(K, grid, s, g, src, cand, self, nth_eff, plane_stride, id_out, d_out,
                              rows, nrows, nrows_dev, reverse, unordered, ctr);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

int32_t launch_brute_exact(const GridView& g, const double* x1, const double* y1, const double* c1,
                           uint32_t n1, const double* x2, const double* y2, const double* c2,
                           uint32_t n2, bool self, uint32_t nth_eff, uint64_t plane_stride,
                           uint64_t* id_out, double* d_out, const uint32_t* rows, uint32_t nrows,
                           bool reverse, SearchCounters* ctr, cudaStream_t s, Error& err,
                           uint64_t* launches) {
    if (nrows == 0 || nth_eff == 0) return XM_OK;
    const dim3 grid((nrows + kBruteThreads - 1)/kBruteThreads);
    const int K = nth_eff <= 8 ? (int)nth_eff : 0;
    if (g.linear) dispatch_brute<true>(K, grid, s, g.sin_model, x1, y1, c1, n1, x2, y2, c2, n2, self,
                                       nth_eff, plane_stride, id_out, d_out, rows, nrows, reverse, ctr);
    else dispatch_brute<false>(K, grid, s, g.sin_model, x1, y1, c1, n1, x2, y2, c2, n2, self, nth_eff,
                               plane_stride, id_out, d_out, rows, nrows, reverse, ctr);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

int32_t launch_prep_unsorted(const double* ra, const double* dec, uint32_t n, bool linear,
                             double* x, double* y, double* c, cudaStream_t s, Error& err,
                             uint64_t* launches, int model) {
    if (n == 0) return XM_OK;
    prep_unsorted_kernel<<<(n + 255)/256, 256, 0, s>>>(ra, dec, n, linear ? 1 : 0, model, x, y, c);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

int32_t launch_fill_outputs(uint64_t* id, double* d, uint64_t n, double dval, cudaStream_t s,
                            Error& err, uint64_t* launches) {
    if (n == 0) return XM_OK;
    uint64_t blocks = (n + 255)/256;
    if (blocks > 148*32) blocks = 148*32;
    fill_outputs_kernel<<<(unsigned)blocks, 256, 0, s>>>(id, d, n, dval);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

int32_t launch_row_tables(const GridView& g, double* cmin_row, double* ccen_row, cudaStream_t s,
                          Error& err, uint64_t* launches) {
    row_tables_kernel<<<(g.ndec + 255)/256, 256, 0, s>>>(g, cmin_row, ccen_row);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

}  // namespace xm
