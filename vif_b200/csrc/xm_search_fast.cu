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
// Work shape: one thread per source, sources and candidates in cell order (32-byte records, one
// LDG.256 each, warp-broadcast through L1 because neighbouring threads sit in the same cells);
// ring 0 is the own cell, ring 1 the other 24 cells of the 5x5 block, visited nearest-first from a
// per-thread survivor bitmask so that lanes do not wait on cells only other lanes need.
#include "xm_common.cuh"
#include "xm_math_exact.cuh"

namespace xm {

namespace {

constexpr int kThreads = 128;

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000ll); }

// 5x5 block offsets, nearest first: 4 sides, 4 corners, 16 outer cells
__constant__ int8_t c_bx[24] = {0, 1, 0, -1,   1, -1, -1, 1,
                                0, 1, 2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1};
__constant__ int8_t c_by[24] = {1, 0, -1, 0,   1, 1, -1, -1,
                                2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1, 0, 1, 2, 2};

__device__ __forceinline__ Rec load_rec(const Rec* p) {
    Rec r;
    unsigned long long w;
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.c), "=l"(w) : "l"(p));
    r.idx = (uint32_t)w;
    r.key = (uint32_t)(w >> 32);
    return r;
}

struct Best {
    double q1, q2;        // smallest and second smallest filter metric
    double x, y, c;       // coordinates of the current winner
    uint32_t idx;
};

__device__ __forceinline__ void scan_cell(const Rec* __restrict__ rec, uint2 r, double x1, double y1,
                                          double c1, uint32_t skip, Best& b, uint32_t& evals) {
    for (uint32_t q = r.x; q < r.y; ++q) {
        const Rec cd = load_rec(rec + q);
        const double dx = cd.x - x1, dy = cd.y - y1;
        double Q = fma(dx*dx*cd.c, c1, dy*dy);
        if (q == skip) Q = dinf();                       // self match: never pair a source with itself
        if (Q < b.q2) {
            if (Q < b.q1) {
                b.q2 = b.q1; b.q1 = Q; b.x = cd.x; b.y = cd.y; b.c = cd.c; b.idx = cd.idx;
            } else if (!(Q == b.q1 && cd.x == b.x && cd.y == b.y)) {
                b.q2 = Q;                                // identical coordinates: exact tie, first stays
            }
        }
    }
    evals += r.y - r.x;
}

// 0: stop for certain, 1: continue for certain, 2: inside the slack
__device__ __forceinline__ int stop_decision(double q1, double tr, const GridView& g) {
    const double lo = tr - g.stop_delta, hi = tr + g.stop_delta;
    if (lo > 0.0 && q1*(1.0 + g.eps_hi) <= lo*lo*(1.0 - lo*lo*(1.0/12.0) - 1e-12)) return 0;
    if (hi <= 0.0) return q1 > 0.0 ? 1 : 2;
    if (q1*(1.0 - g.eps_lo) > hi*hi) return 1;
    return 2;
}

__global__ void __launch_bounds__(kThreads)
ring_nn_filter_kernel(GridView g, CatView src, CatView cand, int self_mode,
                      uint64_t* __restrict__ id_out, double* __restrict__ d_out,
                      uint32_t* __restrict__ refine_list, unsigned long long* refine_count,
                      int reverse, SearchCounters* ctr) {
    const uint32_t p = blockIdx.x*blockDim.x + threadIdx.x;
    uint32_t evals = 0;
    if (p < src.n) {
        const Rec me = load_rec(src.rec + p);
        const double x1 = me.x, y1 = me.y, c1 = me.c;
        const int x0 = (int)(me.key/g.ndec), y0 = (int)(me.key - (uint32_t)x0*g.ndec);
        const uint32_t skip = self_mode ? p : 0xffffffffu;
        const int ndec = (int)g.ndec, nra = (int)g.nra;

        Best b;
        b.q1 = dinf(); b.q2 = dinf(); b.x = 0; b.y = 0; b.c = 0; b.idx = 0xffffffffu;
        bool flag = false, done = false;

        // ---- ring 0: own cell ------------------------------------------------------------------
        scan_cell(cand.rec, cand.cell[me.key], x1, y1, c1, skip, b, evals);

        // distance to the cell centre and the stop rule after ring 0 (max_dist = cs/2)
        const double ex_lo = g.ox + (double)x0*g.wx, ey_lo = g.oy + (double)y0*g.wy;
        const double dxc = (ex_lo + 0.5*g.wx) - x1, dyc = (ey_lo + 0.5*g.wy) - y1;
        const double cdist = sqrt(fma(dxc*dxc*g.ccen_row[y0], c1, dyc*dyc));
        {
            const int dec = stop_decision(b.q1, 0.4*g.cs_rad - 1.1*cdist, g);
            if (dec == 0) done = true; else if (dec == 2) { flag = true; done = true; }
        }

        if (!done) {
            // ---- ring 1: the other 24 cells of the 5x5 block ----------------------------------------
            // squared gaps from the source to the block's cell columns/rows (with the binning slack)
            double gl = x1 - ex_lo - g.fzx, gr = (ex_lo + g.wx) - x1 - g.fzx;
            double gd = y1 - ey_lo - g.fzy, gu = (ey_lo + g.wy) - y1 - g.fzy;
            gl = gl > 0 ? gl : 0; gr = gr > 0 ? gr : 0; gd = gd > 0 ? gd : 0; gu = gu > 0 ? gu : 0;
            const double shrink = 1.0 - g.eps_lo - 1e-9;
            double GX[5], GY[5], CM[5];
            GX[2] = 0.0; GX[3] = gr*gr*c1; GX[4] = (gr + g.wx)*(gr + g.wx)*c1;
            GX[1] = gl*gl*c1; GX[0] = (gl + g.wx)*(gl + g.wx)*c1;
            GY[2] = 0.0; GY[3] = gu*gu; GY[4] = (gu + g.wy)*(gu + g.wy);
            GY[1] = gd*gd; GY[0] = (gd + g.wy)*(gd + g.wy);
#pragma unroll
            for (int k = 0; k < 5; ++k) {
                const int yy = y0 + k - 2;
                CM[k] = (yy >= 0 && yy < ndec) ? g.cmin_row[yy]*shrink : -1.0;
                GY[k] *= shrink;
            }
            // three phases (sides, corners, outer cells): survivors are re-derived from the current
            // best before each phase, then scanned by a per-lane loop over the set bits
#pragma unroll
            for (int phase = 0; phase < 3; ++phase) {
                const int e0 = phase == 0 ? 0 : (phase == 1 ? 4 : 8);
                const int e1 = phase == 0 ? 4 : (phase == 1 ? 8 : 24);
                const double thr = b.q1*(1.0 + g.eps_hi)*(1.0 + 1e-9);
                uint32_t mask = 0;
#pragma unroll
                for (int e = e0; e < e1; ++e) {
                    constexpr int8_t BX[24] = {0, 1, 0, -1, 1, -1, -1, 1,
                                               0, 1, 2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1};
                    constexpr int8_t BY[24] = {1, 0, -1, 0, 1, 1, -1, -1,
                                               2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1, 0, 1, 2, 2};
                    const int bx = BX[e], by = BY[e];
                    const int cx = x0 + bx;
                    const double cm = CM[by + 2];
                    // cm < 0: row outside the grid or beyond a pole -> outside: skip, pole: flag below
                    const double lb = fma(GX[bx + 2], cm, GY[by + 2]);
                    const bool in_grid = cx >= 0 && cx < nra && cm >= 0.0;
                    if (in_grid && lb < thr) mask |= 1u << e;
                }
                while (mask) {
                    const int e = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const uint32_t cell = (uint32_t)(x0 + c_bx[e])*g.ndec + (uint32_t)(y0 + c_by[e]);
                    const uint2 r = cand.cell[cell];
                    if (r.x != r.y) scan_cell(cand.rec, r, x1, y1, c1, skip, b, evals);
                }
            }
            // stop rule after ring 1 (max_dist = 2.5 cs): going on to ring 2 is the exact kernel's job
            const int dec = stop_decision(b.q1, 2.4*g.cs_rad - 1.1*cdist, g);
            if (dec != 0) flag = true;
        }

        // the winner must beat the runner-up by more than the filter's uncertainty
        if (!(b.q1*(1.0 + g.eps_hi) < b.q2*(1.0 - g.eps_lo)) && b.q2 < dinf()) flag = true;
        if (b.idx == 0xffffffffu) flag = true;

        if (flag) {
            const unsigned long long slot = atomicAdd(refine_count, 1ull);
            refine_list[slot] = p;
        } else {
            // exact proxy of the winning pair in reference-order arithmetic -> arcsec
            const double pr = proxy_sph_small(x1, y1, c1, b.x, b.y, b.c, g.sin_model);
            id_out[me.idx] = (uint64_t)b.idx;
            d_out[me.idx] = proxy_to_true(pr, false);
        }
    }
    __syncwarp();
    unsigned long long v = evals;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(reverse ? &ctr->evals_rev : &ctr->evals_fwd, v);
}

}  // namespace

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
                              uint64_t* id_out, double* d_out, uint32_t* refine_list,
                              unsigned long long* refine_count, bool reverse, SearchCounters* ctr,
                              cudaStream_t s, Error& err, uint64_t* launches) {
    if (src.n == 0) return XM_OK;
    ring_nn_filter_kernel<<<(src.n + kThreads - 1)/kThreads, kThreads, 0, s>>>(
        g, src, cand, self ? 1 : 0, id_out, d_out, refine_list, refine_count, reverse ? 1 : 0, ctr);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

}  // namespace xm
