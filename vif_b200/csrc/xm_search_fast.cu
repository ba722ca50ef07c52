// Nearest-neighbour filter kernel for the cell-binned path (nth = 1, spherical): the hot kernel
// of the headline configuration. It decides, for almost every source, which candidate the
// reference's ring search (qxmatch.hpp:293-380) would return, with ~6 FP64 operations per
// candidate instead of two libm sines, and hands the rows it cannot decide *rigorously* to the
// exact kernel (xm_search_exact.cu), which re-does them in reference-order arithmetic.
//
// Filter metric.  With dx = ra2-ra1, dy = dec2-dec1 [rad]:   Q = dy^2 + dx^2*cos(dec1)*cos(dec2).
// The reference's proxy is p = sin^2(dy/2) + sin^2(dx/2)*cos*cos, and h^2(1-h^2/3) <= sin^2 h <= h^2,
// so 4p lies in [Q(1-eps_lo), Q(1+eps_hi)] with eps_lo = H^2/3 + rounding, H the largest half
// difference inside the 5x5 cell block (GridView::eps_lo/eps_hi are computed on the host).
//   * winner: the candidate with the smallest Q wins for certain if Q1*(1+eps_hi) < Q2*(1-eps_lo)
//     (Q2 = runner-up); otherwise the row is flagged. Candidates at bit-identical coordinates are
//     exact ties in the reference too and keep first-arrival order (same cell, ascending row).
//   * culling: a cell whose lower bound exceeds Q1 by the same margin cannot hold a winner nor a
//     near-tie, so it is skipped (the reference would reject all its candidates with strict '<').
//   * stop rule (qxmatch.hpp:336-340): evaluated on the small-angle forms with a slack
//     (GridView::stop_delta); rows whose decision falls inside the slack are flagged. Rows that
//     must go on to ring 2 and beyond are flagged as well (the exact kernel walks any depth).
// Flagged rows are a few per 10^7 on the benchmark catalogs (xm_stats::rows_refined_*).
//
// Work shape: see ring_nn_filter_kernel below (one CTA = a strip of one grid column, candidates
// staged in shared memory by the TMA engine, ring-1 work redistributed over the CTA).
#include "xm_common.cuh"
#include "xm_math_exact.cuh"
#include "xm_filter_common.cuh"

namespace xm {

namespace {

using namespace filt;

constexpr int kThreads = 256;

struct Best {
    int32_t h1, h2;
    uint32_t slot;
};

template <bool SELF>
__device__ __forceinline__ void eval_loaded(const Rec& cd, uint32_t q, double x1, double y1, double c1,
                                            uint32_t self_slot, Best& b) {
    const double dx = cd.x - x1, dy = cd.y - y1;
    const double Q = fma(dx*dx*cd.c, c1, dy*dy);
    int32_t h = __double2hiint(Q);
    if (SELF && q == self_slot) h = kInfHi;             // never pair a source with itself
    const bool better = h < b.h1;
    b.h2 = min(b.h2, max(b.h1, h));
    b.h1 = min(b.h1, h);
    b.slot = better ? q : b.slot;
}

// all candidates of one cell read from global memory; four independent 32-byte loads in flight
template <bool SELF>
__device__ __forceinline__ void scan_range(const Rec* __restrict__ rec, uint32_t qs, uint32_t qe,
                                           double x1, double y1, double c1, uint32_t self_slot,
                                           Best& b) {
    uint32_t q = qs;
    for (; q + 4 <= qe; q += 4) {
        const Rec r0 = load_rec(rec + q), r1 = load_rec(rec + q + 1);
        const Rec r2 = load_rec(rec + q + 2), r3 = load_rec(rec + q + 3);
        eval_loaded<SELF>(r0, q, x1, y1, c1, self_slot, b);
        eval_loaded<SELF>(r1, q + 1, x1, y1, c1, self_slot, b);
        eval_loaded<SELF>(r2, q + 2, x1, y1, c1, self_slot, b);
        eval_loaded<SELF>(r3, q + 3, x1, y1, c1, self_slot, b);
    }
    if (q + 2 <= qe) {
        const Rec r0 = load_rec(rec + q), r1 = load_rec(rec + q + 1);
        eval_loaded<SELF>(r0, q, x1, y1, c1, self_slot, b);
        eval_loaded<SELF>(r1, q + 1, x1, y1, c1, self_slot, b);
        q += 2;
    }
    if (q < qe) {
        const Rec r0 = load_rec(rec + q);
        eval_loaded<SELF>(r0, q, x1, y1, c1, self_slot, b);
    }
}

// the same cell read from the CTA's staged copy: slot q lives at stage[q + delta]
template <bool SELF>
__device__ __forceinline__ void scan_range_staged(const Rec* stage, int32_t delta, uint32_t qs,
                                                  uint32_t qe, double x1, double y1, double c1,
                                                  uint32_t self_slot, Best& b) {
#pragma unroll 2
    for (uint32_t q = qs; q < qe; ++q) eval_loaded<SELF>(stage[(int32_t)q + delta], q, x1, y1, c1, self_slot, b);
}

// merge the (smallest, second smallest, slot) triple of another candidate set into b
__device__ __forceinline__ void merge_best(Best& b, int32_t a1, int32_t a2, uint32_t aslot) {
    const bool better = a1 < b.h1;
    b.h2 = min(min(b.h2, a2), max(b.h1, a1));
    b.h1 = min(b.h1, a1);
    b.slot = better ? aslot : b.slot;
}

constexpr int kQueuePerSrc = 2;          // ring-1 cells a source may queue; further ones are scanned in place
constexpr int kMaxItems = kQueuePerSrc*kThreads;   // typical total: ~0.7 per source
constexpr int kStageCap = 1024;          // candidate records a CTA can stage (32 KB; typical need ~800)
constexpr int kStageRows = 64;           // cell rows (incl. halo) a staged block may span

// Per-tile staging plan, written by plan_tiles_kernel once per search: where the three column runs
// of the tile's neighbourhood start in the cell-ordered candidate array.
struct __align__(16) TilePlan {
    uint32_t run_start[3];
    uint32_t run_cnt[3];
    int32_t  ya;                 // first staged cell row
    uint32_t xs_nrows;           // column of the strip | staged rows << 24 (0 rows: not staged)
};

// dynamic shared memory layout of ring_nn_filter_kernel
struct FilterSmem {
    Rec      cand[kStageCap];
    double   x[kThreads], y[kThreads], c[kThreads];
    uint2    range[kMaxItems];
    int32_t  shift[kMaxItems];
    int32_t  a1[kMaxItems], a2[kMaxItems];
    uint32_t slot[kMaxItems];
    uint32_t cs[3][kStageRows + 1];
    uint64_t bar;
    uint32_t warp_tot[kThreads/32];
    uint8_t  owner[kMaxItems];
};

__global__ void __launch_bounds__(256)
plan_tiles_kernel(GridView g, CatView src, CatView cand, uint32_t ntiles, TilePlan* __restrict__ plan) {
    const uint32_t t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= ntiles) return;
    const uint32_t p0 = t*kThreads, p1 = min(p0 + (uint32_t)kThreads, src.n) - 1;
    const uint32_t kf = src.rec[p0].key, kl = src.rec[p1].key;
    const int ndec = (int)g.ndec, nra = (int)g.nra;
    const int xs = (int)(kf/g.ndec), yf = (int)(kf - (uint32_t)xs*g.ndec);
    const int xl = (int)(kl/g.ndec), yl = (int)(kl - (uint32_t)xl*g.ndec);
    const int ya = max(yf - 1, 0), yb = min(yl + 1, ndec - 1);
    const int nrows = yb - ya + 1;
    TilePlan pl;
    pl.ya = ya; pl.xs_nrows = 0;
    for (int c = 0; c < 3; ++c) { pl.run_start[c] = 0; pl.run_cnt[c] = 0; }
    if (xs == xl && nrows <= kStageRows && xs < (1 << 24)) {
        uint32_t tot = 0;
        for (int c = 0; c < 3; ++c) {
            const int cx = xs - 1 + c;
            if (cx < 0 || cx >= nra) continue;
            const uint32_t b0 = cand.cstart[(uint32_t)cx*g.ndec + (uint32_t)ya];
            const uint32_t b1 = cand.cstart[(uint32_t)cx*g.ndec + (uint32_t)(ya + nrows)];
            pl.run_start[c] = b0; pl.run_cnt[c] = b1 - b0;
            tot += b1 - b0;
        }
        if (tot <= (uint32_t)kStageCap) pl.xs_nrows = (uint32_t)xs | ((uint32_t)nrows << 24);
    }
    plan[t] = pl;
}

constexpr int32_t kNotStaged = INT32_MIN;

// ---- TMA bulk copy (1-D, no tensor map) + mbarrier ------------------------------------------------
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

// Work shape of one CTA (128 consecutive sources in cell order = a short strip of one grid column):
//   0. the candidate records of the strip's 3-column neighbourhood (rows +-1) are contiguous slot
//      runs of the cell-ordered catalog (one per column): the TMA engine copies them into shared
//      memory (cp.async.bulk + mbarrier) together with their CSR offsets; strips that straddle two
//      columns or overflow the stage read global memory instead (same code path otherwise);
//   1. every thread scans its source's own cell (ring 0), applies the stop rule and derives the
//      bitmask of ring-1 cells that survive the culling test;
//   2. the surviving non-empty (source, cell) pairs are queued in shared memory (block prefix sum);
//   3. the CTA's threads take one queued pair each -- a source that needs three neighbouring cells
//      no longer stalls the 31 lanes next to it that need none;
//   4. every thread merges the partial results of its own source, applies the stop rule of ring 1
//      and the ambiguity test, and writes id / d (or queues the row for the exact kernel).
template <bool SELF>
__global__ void __launch_bounds__(kThreads, 4)
ring_nn_filter_kernel(GridView g, CatView src, CatView cand, const TilePlan* __restrict__ plan,
                      ResRec* __restrict__ res, uint32_t* __restrict__ refine_list,
                      unsigned long long* refine_count, int reverse, SearchCounters* ctr) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FilterSmem& sm = *reinterpret_cast<FilterSmem*>(smem_raw);
    Rec* const s_cand = sm.cand;

    const uint32_t p0 = blockIdx.x*blockDim.x;
    const uint32_t p = p0 + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool live = p < src.n;
    const int ndec = (int)g.ndec, nra = (int)g.nra;
    uint32_t evals = 0;

    // ---- 0. stage the neighbourhood (plan and own record are loaded together) ---------------------
    const uint4 pa = __ldg(reinterpret_cast<const uint4*>(plan + blockIdx.x));
    const uint4 pb = __ldg(reinterpret_cast<const uint4*>(plan + blockIdx.x) + 1);
    Rec me;
    me.x = me.y = me.c = 0.0; me.idx = 0; me.key = 0;
    if (live) me = load_rec(src.rec + p);
    if (threadIdx.x == 0) mbar_init(&sm.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();

    const uint32_t run_start[3] = {pa.x, pa.y, pa.z};
    const uint32_t run_cnt[3] = {pa.w, pb.x, pb.y};
    const int ya = (int)pb.z;
    const int nrows = (int)(pb.w >> 24);
    const int xs = (int)(pb.w & 0xffffffu);
    const bool staged_plan = nrows > 0;
    const uint32_t tot = run_cnt[0] + run_cnt[1] + run_cnt[2];
    int32_t delta[3];
    delta[0] = -(int32_t)run_start[0];
    delta[1] = (int32_t)run_cnt[0] - (int32_t)run_start[1];
    delta[2] = (int32_t)(run_cnt[0] + run_cnt[1]) - (int32_t)run_start[2];
    if (staged_plan) {
        if (threadIdx.x == 0 && tot) {
            mbar_expect_tx(&sm.bar, tot*(uint32_t)sizeof(Rec));
            uint32_t off = 0;
#pragma unroll
            for (int col = 0; col < 3; ++col) {
                if (run_cnt[col]) bulk_g2s(&s_cand[off], cand.rec + run_start[col], run_cnt[col]*(uint32_t)sizeof(Rec), &sm.bar);
                off += run_cnt[col];
            }
        }
        for (int t = threadIdx.x; t < 3*(nrows + 1); t += kThreads) {
            const int col = t/(nrows + 1), r = t - col*(nrows + 1);
            const int cx = xs - 1 + col;
            sm.cs[col][r] = (cx >= 0 && cx < nra) ? cand.cstart[(uint32_t)cx*g.ndec + (uint32_t)(ya + r)] : 0u;
        }
    }
    __syncthreads();
    const int staged = staged_plan ? 1 : 0;
    if (staged && tot) mbar_wait(&sm.bar, 0);

    Best b;
    b.h1 = kInfHi; b.h2 = kInfHi; b.slot = 0xffffffffu;
    bool flag = false, done = !live;
    uint32_t mask = 0;
    int x0 = 0, y0 = 0;
    double cdist = 0.0;

    if (live) {
        if (staged) x0 = xs; else x0 = (int)(me.key/g.ndec);       // one-column strips skip the division
        y0 = (int)(me.key - (uint32_t)x0*g.ndec);
        const double x1 = me.x, y1 = me.y, c1 = me.c;
        // ---- 1. ring 0: own cell ---------------------------------------------------------------------
        if (staged) {
            const uint32_t qs = sm.cs[1][y0 - ya], qe = sm.cs[1][y0 - ya + 1];
            scan_range_staged<SELF>(s_cand, delta[1], qs, qe, x1, y1, c1, p, b);
            evals += qe - qs;
        } else {
            const uint32_t qs = cand.cstart[me.key], qe = cand.cstart[me.key + 1];
            scan_range<SELF>(cand.rec, qs, qe, x1, y1, c1, p, b);
            evals += qe - qs;
        }
        // distance to the cell centre and the stop rule after ring 0 (max_dist = cs/2)
        const double ex_lo = g.ox + (double)x0*g.wx, ey_lo = g.oy + (double)y0*g.wy;
        const double dxc = (ex_lo + 0.5*g.wx) - x1, dyc = (ey_lo + 0.5*g.wy) - y1;
        cdist = sqrt(fma(dxc*dxc*g.ccen_row[y0], c1, dyc*dyc));
        {
            const int dec = stop_decision(hi_floor(b.h1), hi_ceil(b.h1), 0.4*g.cs_rad - 1.1*cdist, g);
            if (dec == 0) done = true; else if (dec == 2) { flag = true; done = true; }
        }
        if (!done) {
            // ---- ring 1: which of the other 24 cells of the 5x5 block can hold a winner? ---------
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
            const double thr = hi_ceil(b.h1)*(1.0 + g.eps_hi)*(1.0 + 1e-9);
            // the 16 outer cells lie at least one full cell away: one test rules them all out for
            // most sources (min over the four directions of the bound of the nearest outer cell)
            const double cm_lo = fmin(fmin(CM[0], CM[1]), fmin(fmin(CM[2], CM[3]), CM[4]));
            const bool outer = !(fmin(fmin(GX[0], GX[4])*fmax(cm_lo, 0.0), fmin(GY[0], GY[4])) >= thr) || cm_lo < 0.0;
#pragma unroll
            for (int e = 0; e < 24; ++e) {
                constexpr int8_t BX[24] = {0, 1, 0, -1, 1, -1, -1, 1,
                                           0, 1, 2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1};
                constexpr int8_t BY[24] = {1, 0, -1, 0, 1, 1, -1, -1,
                                           2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1, 0, 1, 2, 2};
                if (e == 8 && !outer) break;
                const int bx = BX[e], by = BY[e];
                const int cx = x0 + bx;
                const double cm = CM[by + 2];
                const double lb = fma(GX[bx + 2], cm, GY[by + 2]);
                if (cx >= 0 && cx < nra && cm >= 0.0 && lb < thr) mask |= 1u << e;
            }
        }
    }

    // ---- 2. queue the surviving (source, cell) pairs ---------------------------------------------------
    // The owner resolves each cell's slot range (from the staged CSR offsets when the cell is inside
    // the staged block) and queues only non-empty cells.
    sm.x[threadIdx.x] = me.x; sm.y[threadIdx.x] = me.y; sm.c[threadIdx.x] = me.c;
    uint2 rng[kQueuePerSrc];
    int32_t shf[kQueuePerSrc];
    uint32_t cnt = 0;
    {
        uint32_t m = mask;
        while (m) {
            const int e = __ffs(m) - 1;
            m &= m - 1;
            const int bx = c_bx[e], by = c_by[e];
            const int ry = y0 + by - ya;
            uint2 r; int32_t sh;
            if (staged && bx >= -1 && bx <= 1 && ry >= 0 && ry < nrows) {
                r = make_uint2(sm.cs[bx + 1][ry], sm.cs[bx + 1][ry + 1]);
                sh = bx < 0 ? delta[0] : (bx > 0 ? delta[2] : delta[1]);
            } else {
                const uint32_t ck = (uint32_t)((int)me.key + bx*ndec + by);
                r = make_uint2(cand.cstart[ck], cand.cstart[ck + 1]);
                sh = kNotStaged;
            }
            if (r.x == r.y) continue;
            if (cnt < (uint32_t)kQueuePerSrc) {
                // static indexing keeps rng[] / shf[] in registers
                if (cnt == 0) { rng[0] = r; shf[0] = sh; } else { rng[1] = r; shf[1] = sh; }
                ++cnt;
            } else {
                // more cells than a source may queue: scan the rest in place
                if (sh != kNotStaged) scan_range_staged<SELF>(s_cand, sh, r.x, r.y, me.x, me.y, me.c, p, b);
                else scan_range<SELF>(cand.rec, r.x, r.y, me.x, me.y, me.c, p, b);
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
    // total <= kMaxItems by construction
#pragma unroll
    for (int k = 0; k < kQueuePerSrc; ++k) {
        if ((uint32_t)k < cnt) {
            sm.owner[first + k] = (uint8_t)threadIdx.x; sm.range[first + k] = rng[k]; sm.shift[first + k] = shf[k];
        }
    }
    __syncthreads();

    // ---- 3. one queued pair per thread ------------------------------------------------------------------
    for (uint32_t it = threadIdx.x; it < total; it += kThreads) {
        const int owner = sm.owner[it];
        const uint2 r = sm.range[it];
        const int32_t sh = sm.shift[it];
        Best a;
        a.h1 = kInfHi; a.h2 = kInfHi; a.slot = 0xffffffffu;
        if (sh != kNotStaged) scan_range_staged<SELF>(s_cand, sh, r.x, r.y, sm.x[owner], sm.y[owner], sm.c[owner], p0 + owner, a);
        else scan_range<SELF>(cand.rec, r.x, r.y, sm.x[owner], sm.y[owner], sm.c[owner], p0 + owner, a);
        evals += r.y - r.x;
        sm.a1[it] = a.h1; sm.a2[it] = a.h2; sm.slot[it] = a.slot;
    }
    __syncthreads();
    for (uint32_t k = 0; k < cnt; ++k) merge_best(b, sm.a1[first + k], sm.a2[first + k], sm.slot[first + k]);

    // ---- 4. decide ----------------------------------------------------------------------------------------
    if (live) {
        if (!done) {
            // stop rule after ring 1 (max_dist = 2.5 cs): going on to ring 2 is the exact kernel's job
            const int dec = stop_decision(hi_floor(b.h1), hi_ceil(b.h1), 2.4*g.cs_rad - 1.1*cdist, g);
            if (dec != 0) flag = true;
        }
        // the winner must beat the runner-up by more than the filter's uncertainty (including the
        // truncation to high words)
        if (b.h2 < kInfHi && !(hi_ceil(b.h1)*(1.0 + g.eps_hi) < hi_floor(b.h2)*(1.0 - g.eps_lo))) flag = true;
        if (b.slot == 0xffffffffu) flag = true;

        if (flag) {
            const unsigned long long pos = atomicAdd(refine_count, 1ull);
            refine_list[pos] = p;
            store_res(res + me.idx, kUndecidedId, 0.0);
        } else {
            // exact proxy of the winning pair in reference-order arithmetic -> arcsec
            const Rec w = load_rec(cand.rec + b.slot);
            const double pr = proxy_sph_small(me.x, me.y, me.c, w.x, w.y, w.c, g.sin_model);
            // one full 32-byte sector per source at its original row: a scattered 8-byte store would
            // cost a sector read-modify-write in DRAM for each of id and d
            store_res(res + me.idx, (uint64_t)w.idx, proxy_to_true_small(pr));
        }
    }
    unsigned long long v = evals;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && v) atomicAdd(reverse ? &ctr->evals_rev : &ctr->evals_fwd, v);
}

__global__ void __launch_bounds__(256)
unpack_results_kernel(const ResRec* __restrict__ res, uint32_t n, uint64_t* __restrict__ id_out,
                      double* __restrict__ d_out) {
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long a, b, c, d;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(res + i));
    if (a == kUndecidedId) return;          // the exact kernel owns this row
    id_out[i] = a;
    d_out[i] = __longlong_as_double((long long)b);
}

}  // namespace

size_t filter_result_bytes(uint64_t n) { return (size_t)n*sizeof(ResRec); }

int32_t launch_unpack_results(const void* res, uint32_t n, uint64_t* id_out, double* d_out, cudaStream_t s,
                              Error& err, uint64_t* launches) {
    if (n == 0) return XM_OK;
    unpack_results_kernel<<<(n + 255)/256, 256, 0, s>>>((const ResRec*)res, n, id_out, d_out);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

// The filter's bounds assume small cells well away from the poles; everything else is the exact
// kernel's job.
bool fast_path_applicable(const GridView& g) {
    if (g.linear) return false;
    if (!(g.wx > 0 && g.wy > 0) || g.wx > 5e-3 || g.wy > 5e-3) return false;   // cells <= 0.29 deg
    const double half_pi = 1.5707963267948966;
    const double lo = g.oy, hi = g.oy + (double)g.ndec*g.wy;
    if (lo < -half_pi || hi > half_pi) return false;
    return true;
}

int32_t launch_ring_nn_filter(const GridView& g, const CatView& src, const CatView& cand, bool self,
                              uint64_t* id_out, double* d_out, void* result_scratch, void* plan_scratch,
                              uint32_t* refine_list, unsigned long long* refine_count, bool reverse,
                              SearchCounters* ctr, cudaStream_t s, cudaEvent_t filter_done, Error& err,
                              uint64_t* launches) {
    if (src.n == 0) return XM_OK;
    const uint32_t ntiles = (src.n + kThreads - 1)/kThreads;
    ResRec* res = (ResRec*)result_scratch;
    TilePlan* plan = (TilePlan*)plan_scratch;
    // > 48 KB of dynamic shared memory is an opt-in, per device
    if (self) XM_CUDA_TRY(cudaFuncSetAttribute(ring_nn_filter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)sizeof(FilterSmem)));
    else XM_CUDA_TRY(cudaFuncSetAttribute(ring_nn_filter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)sizeof(FilterSmem)));
    plan_tiles_kernel<<<(ntiles + 255)/256, 256, 0, s>>>(g, src, cand, ntiles, plan);
    if (self) ring_nn_filter_kernel<true><<<ntiles, kThreads, sizeof(FilterSmem), s>>>(
        g, src, cand, plan, res, refine_list, refine_count, reverse ? 1 : 0, ctr);
    else ring_nn_filter_kernel<false><<<ntiles, kThreads, sizeof(FilterSmem), s>>>(
        g, src, cand, plan, res, refine_list, refine_count, reverse ? 1 : 0, ctr);
    // the list of undecided rows is complete here: the caller starts the exact kernel on a side
    // stream from this event, concurrently with the unpack pass (which skips those rows)
    if (filter_done) XM_CUDA_TRY(cudaEventRecord(filter_done, s));
    unpack_results_kernel<<<(src.n + 255)/256, 256, 0, s>>>(res, src.n, id_out, d_out);
    *launches += 3;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

size_t filter_plan_bytes(uint64_t n) { return (size_t)((n + kThreads - 1)/kThreads)*sizeof(TilePlan); }

}  // namespace xm
