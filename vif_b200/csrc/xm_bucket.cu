// Cell index by direct bucket placement -- the fast path of the "bucket fill" (qxmatch.hpp:270-288).
//
// The reference pushes row ids into per-cell vectors in ascending row order. Rebuilt here as a
// counting sort whose every pass streams (no random 8-byte gathers, only full-sector scatters):
//   1. count:  one pass over (ra, dec): cell key (IEEE subtract/divide/floor as the reference),
//              atomic add on the per-cell counter (the table is L2-resident; a reduction: nothing returned);
//   2. scan:   exclusive prefix sum of the counters -> CSR offsets cstart[ncell+1];
//   3. place:  second pass over (ra, dec): the slot is drawn from a cursor that starts as a copy of cstart
//              (atomicAdd); the source's 32-byte record (radians, cos(dec), row, key) is written to that slot -- a
//              scattered store of one full DRAM sector, no read-modify-write;
//   4. order (skipped for nearest-neighbour-only calls, whose kernels do not depend on the order
//              inside a cell):  atomics fill a cell in arbitrary order, the reference's order inside a cell is
//              ascending row (it decides exact ties): one warp per cell ranks the cell's records by
//              row and writes them to their final slots (out of place, both sides streamed).
// Cells holding more than kMaxCellSort rows are left unordered and reported through *overflow; the
// caller then rebuilds that catalog with the stable radix sort (xm_sort.cu), which has no such limit.
#include <cstdlib>

#include "xm_common.cuh"
#include "xm_math_exact.cuh"

namespace xm {

namespace {

constexpr int kThreads = 256;
constexpr uint32_t kMaxCellSort = 4096;

// on_grid: the row is finite and falls on the grid. With bounding boxes measured by the library that is always so
// (the grid is padded by a cell, qxmatch.hpp:258-262); with boxes promised by the caller it is the check that the
// promise held (rows are clamped onto the grid so that nothing is written out of bounds, and the call fails).
__device__ __forceinline__ uint32_t cell_key(double ra, double dec, const GridView& g, bool* on_grid = nullptr) {
    // qxmatch.hpp:278-279
    const double fx = floor(__ddiv_rn(__dsub_rn(ra, g.rra0), g.dra));
    const double fy = floor(__ddiv_rn(__dsub_rn(dec, g.rdec0), g.ddec));
    if (on_grid) *on_grid = fx >= 0 && fx < (double)g.nra && fy >= 0 && fy < (double)g.ndec;     // false for NaN
    // clamped onto the grid; written so that a NaN takes the first branch (a float -> integer conversion of NaN
    // is not something to rely on)
    const uint32_t ix = !(fx >= 0) ? 0u : (fx >= (double)g.nra ? g.nra - 1 : (uint32_t)fx);
    const uint32_t iy = !(fy >= 0) ? 0u : (fy >= (double)g.ndec ? g.ndec - 1 : (uint32_t)fy);
    return ix*g.ndec + iy;
}

constexpr int kRows = 4;     // rows per thread: independent loads / atomics in flight

// ---- exclusive scan of u32 (three-level, in place) ------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads*kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    uint32_t off = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < kScanThreads/32; ++k) { if (k < w) off += warp_sums[k]; tot += warp_sums[k]; }
    total = tot;
    __syncthreads();
    return off + inc - v;
}

// each CTA scans its tile locally and emits the tile total
__global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(uint32_t* __restrict__ data, uint32_t n, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t ws[kScanThreads/32];
    const uint32_t base = blockIdx.x*kScanTile + threadIdx.x*kScanItems;
    uint32_t v[kScanItems], sum = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { v[k] = base + k < n ? data[base + k] : 0u; sum += v[k]; }
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, ws, total);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { if (base + k < n) data[base + k] = run; run += v[k]; }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads)
scan_add_kernel(uint32_t* __restrict__ data, uint32_t n, const uint32_t* __restrict__ tile_offs) {
    const uint32_t off = tile_offs[blockIdx.x];
    const uint32_t base = blockIdx.x*kScanTile + threadIdx.x*kScanItems;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) if (base + k < n) data[base + k] += off;
}

// in-place exclusive scan of data[0..n); aux must hold ceil(n/tile) + ceil(n/tile^2) + 2 words
void exclusive_scan(uint32_t* data, uint32_t n, uint32_t* aux, cudaStream_t s, uint64_t* launches) {
    const uint32_t t1 = (n + kScanTile - 1)/kScanTile;
    scan_tiles_kernel<<<t1, kScanThreads, 0, s>>>(data, n, aux);
    *launches += 1;
    if (t1 > 1) {
        exclusive_scan(aux, t1, aux + t1, s, launches);
        scan_add_kernel<<<t1, kScanThreads, 0, s>>>(data, n, aux);
        *launches += 1;
    }
}

// ---- count and placement --------------------------------------------------------------------------------------
// pass 1: per-cell counters. The pass only adds (no value returned: a reduction, which does not hold the warp until
// the atomic comes back; round 1 kept the returned value as the row's arrival rank: 98 us against 64 us per 1e7 rows).
__global__ void __launch_bounds__(kThreads)
count_kernel(const double* __restrict__ ra, const double* __restrict__ dec, uint32_t n, GridView g,
                 uint32_t* __restrict__ count) {
    const uint32_t base = blockIdx.x*(kThreads*kRows) + threadIdx.x;
    double r[kRows], d[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const uint32_t i = base + k*kThreads;
        r[k] = i < n ? __ldg(ra + i) : 0.0;
        d[k] = i < n ? __ldg(dec + i) : 0.0;
    }
    uint32_t off = 0;
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const uint32_t i = base + k*kThreads;
        if (i >= n) continue;
        bool on = true;
        atomicAdd(count + cell_key(r[k], d[k], g, &on), 1u);
        off += on ? 0u : 1u;
    }
    if (off && g.off_grid) atomicAdd(g.off_grid, (unsigned long long)off);
}

// pass 2: the row's slot is drawn from a cursor that starts as a copy of the CSR offsets (atomicAdd, L2-resident);
// one full-sector store per source
__global__ void __launch_bounds__(kThreads)
place_kernel(const double* __restrict__ ra, const double* __restrict__ dec, uint32_t n, GridView g,
             uint32_t* __restrict__ cursor, Rec* __restrict__ out) {
    const uint32_t base = blockIdx.x*(kThreads*kRows) + threadIdx.x;
    double r[kRows], d[kRows];
    uint32_t key[kRows], slot[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const uint32_t i = base + k*kThreads;
        r[k] = i < n ? __ldg(ra + i) : 0.0;
        d[k] = i < n ? __ldg(dec + i) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const uint32_t i = base + k*kThreads;
        bool on = true;
        key[k] = cell_key(r[k], d[k], g, &on);
        slot[k] = i < n ? atomicAdd(cursor + key[k], 1u) : 0u;
        if (!on && g.off_grid) {
            // cell-binned path only (the spherical grid of xm_sphere.cu clamps its edge rows on purpose and passes no
            // counter): a row that is not finite or off the grid (promised bounding boxes that did not hold: the call
            // fails when it ends) is moved to the centre of the cell it was clamped to, so that every kernel
            // downstream sees coordinates that agree with the index
            r[k] = g.rra0 + ((double)(key[k]/g.ndec) + 0.5)*g.dra;
            d[k] = g.rdec0 + ((double)(key[k]%g.ndec) + 0.5)*g.ddec;
        }
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const uint32_t i = base + k*kThreads;
        if (i >= n) continue;
        double x, y, c;
        if (g.linear) { x = r[k]; y = d[k]; c = 1.0; }
        else { y = __dmul_rn(d[k], XM_D2R); x = __dmul_rn(r[k], XM_D2R); c = cos_ref(y, g.sin_model); }   // qxmatch.hpp:174-180
        const unsigned long long w = (unsigned long long)i | ((unsigned long long)key[k] << 32);
        asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" :: "l"(out + slot[k]), "l"(__double_as_longlong(x)),
                     "l"(__double_as_longlong(y)), "l"(__double_as_longlong(c)), "l"(w) : "memory");
    }
}

// ---- the same two passes for cell tables that do not fit the L2 cache ---------------------------------------------
// (the spherical grid of a 1e9-row catalog has 1e8 cells: a returning atomic on a DRAM-resident cursor costs more than
// the plain read of cstart[key] it replaces, and the cursor is another 4 bytes per cell to copy: C5 on one GPU
// 174 -> 181 ms with the cursor form). Here the counting atomic's return value is kept as the row's arrival rank and
// the placement needs no atomic.
__global__ void __launch_bounds__(kThreads)
count_rank_kernel(const double* __restrict__ ra, const double* __restrict__ dec, uint32_t n, GridView g,
             uint32_t* __restrict__ count, uint32_t* __restrict__ rank) {
    const uint32_t base = blockIdx.x*(kThreads*kRows) + threadIdx.x;
    double r[kRows], d[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const uint32_t i = base + k*kThreads;
        r[k] = i < n ? __ldg(ra + i) : 0.0;
        d[k] = i < n ? __ldg(dec + i) : 0.0;
    }
    uint32_t rk[kRows];
    uint32_t off = 0;
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const uint32_t i = base + k*kThreads;
        bool on = true;
        rk[k] = i < n ? atomicAdd(count + cell_key(r[k], d[k], g, &on), 1u) : 0u;
        off += (i < n && !on) ? 1u : 0u;
    }
    if (off && g.off_grid) atomicAdd(g.off_grid, (unsigned long long)off);
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const uint32_t i = base + k*kThreads;
        if (i < n) rank[i] = rk[k];
    }
}

// pass 2: slot = cstart[key] + arrival rank; one full-sector store per source, no atomics
__global__ void __launch_bounds__(kThreads)
place_rank_kernel(const double* __restrict__ ra, const double* __restrict__ dec, uint32_t n, GridView g,
             const uint32_t* __restrict__ cstart, const uint32_t* __restrict__ rank, Rec* __restrict__ out) {
    const uint32_t base = blockIdx.x*(kThreads*kRows) + threadIdx.x;
    double r[kRows], d[kRows];
    uint32_t rk[kRows], key[kRows], slot[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const uint32_t i = base + k*kThreads;
        r[k] = i < n ? __ldg(ra + i) : 0.0;
        d[k] = i < n ? __ldg(dec + i) : 0.0;
        rk[k] = i < n ? __ldg(rank + i) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        bool on = true;
        key[k] = cell_key(r[k], d[k], g, &on);
        slot[k] = __ldg(cstart + key[k]) + rk[k];
        if (!on && g.off_grid) {           // cell-binned path only: see place_kernel
            r[k] = g.rra0 + ((double)(key[k]/g.ndec) + 0.5)*g.dra;
            d[k] = g.rdec0 + ((double)(key[k]%g.ndec) + 0.5)*g.ddec;
        }
    }
#pragma unroll
    for (int k = 0; k < kRows; ++k) {
        const uint32_t i = base + k*kThreads;
        if (i >= n) continue;
        double x, y, c;
        if (g.linear) { x = r[k]; y = d[k]; c = 1.0; }
        else { y = __dmul_rn(d[k], XM_D2R); x = __dmul_rn(r[k], XM_D2R); c = cos_ref(y, g.sin_model); }   // qxmatch.hpp:174-180
        const unsigned long long w = (unsigned long long)i | ((unsigned long long)key[k] << 32);
        asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" :: "l"(out + slot[k]), "l"(__double_as_longlong(x)),
                     "l"(__double_as_longlong(y)), "l"(__double_as_longlong(c)), "l"(w) : "memory");
    }
}

// one warp per cell: rank the cell's records by original row, write them to their final slots
__global__ void __launch_bounds__(kThreads)
order_cells_kernel(const Rec* __restrict__ in, const uint32_t* __restrict__ cstart, uint32_t ncell,
                   Rec* __restrict__ out, uint32_t* __restrict__ overflow) {
    const uint32_t warp = (blockIdx.x*blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= ncell) return;
    const uint32_t s = cstart[warp], m = cstart[warp + 1] - s;
    if (m == 0) return;
    if (m <= 32) {
        Rec r;
        r.x = r.y = r.c = 0.0; r.idx = 0xffffffffu; r.key = 0;
        if ((uint32_t)lane < m) r = in[s + lane];
        uint32_t rank = 0;
        for (uint32_t j = 0; j < m; ++j) rank += __shfl_sync(0xffffffffu, r.idx, (int)j) < r.idx ? 1u : 0u;
        if ((uint32_t)lane < m) out[s + rank] = r;
    } else if (m <= kMaxCellSort) {
        for (uint32_t e = lane; e < m; e += 32) {
            const Rec r = in[s + e];
            uint32_t rank = 0;
            for (uint32_t f = 0; f < m; ++f) rank += in[s + f].idx < r.idx ? 1u : 0u;
            out[s + rank] = r;
        }
    } else {
        for (uint32_t e = lane; e < m; e += 32) out[s + e] = in[s + e];
        if (lane == 0) atomicAdd(overflow, 1u);
    }
}

// ---- mutual-best filter (xmatch_clean_best, qxmatch.hpp:641-663) ------------------------------------
__global__ void __launch_bounds__(kThreads)
clean_flags_kernel(const uint64_t* __restrict__ id, const uint64_t* __restrict__ rid, uint32_t n1,
                   uint64_t n2, uint32_t* __restrict__ keep) {
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i > n1) return;
    uint32_t k = 0;
    if (i < n1) {
        const uint64_t j = id[i];                      // rank-0 plane: nearest neighbour of row i
        k = (j < n2 && rid[j] == (uint64_t)i) ? 1u : 0u;
    }
    keep[i] = k;                                       // keep[n1] = 0: the scan leaves the total there
}

__global__ void __launch_bounds__(kThreads)
clean_compact_kernel(const uint64_t* __restrict__ id, const uint32_t* __restrict__ pos, uint32_t n1,
                     uint64_t* __restrict__ id1, uint64_t* __restrict__ id2, uint64_t* __restrict__ lost,
                     uint64_t* __restrict__ counts) {
    const uint32_t i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= n1) return;
    const uint32_t p = pos[i], kept = pos[i + 1] - p;
    if (kept) { id1[p] = i; id2[p] = id[i]; }
    else lost[i - p] = i;
    if (i == 0) { counts[0] = pos[n1]; counts[1] = (uint64_t)n1 - pos[n1]; }
}

}  // namespace

size_t clean_best_scratch_bytes(uint32_t n1) { return ((size_t)n1 + 1)*sizeof(uint32_t) + bucket_scan_aux_bytes(n1); }

// id1/id2: rows of catalog 1 / 2 forming mutual best pairs (ascending row of catalog 1); lost: the
// other rows of catalog 1; counts = {number of pairs, number lost}. scratch: clean_best_scratch_bytes.
int32_t launch_clean_best(const uint64_t* id, const uint64_t* rid, uint32_t n1, uint64_t n2, uint64_t* id1,
                          uint64_t* id2, uint64_t* lost, uint64_t* counts, void* scratch, cudaStream_t s,
                          Error& err, uint64_t* launches) {
    uint32_t* pos = (uint32_t*)scratch;
    uint32_t* aux = pos + (size_t)n1 + 1;
    if (n1 == 0) { XM_CUDA_TRY(cudaMemsetAsync(counts, 0, 2*sizeof(uint64_t), s)); return XM_OK; }
    clean_flags_kernel<<<(n1 + 1 + kThreads - 1)/kThreads, kThreads, 0, s>>>(id, rid, n1, n2, pos);
    *launches += 1;
    exclusive_scan(pos, n1 + 1, aux, s, launches);
    clean_compact_kernel<<<(n1 + kThreads - 1)/kThreads, kThreads, 0, s>>>(id, pos, n1, id1, id2, lost, counts);
    *launches += 1;
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

// in-place exclusive scan of data[0..n) for other translation units (aux: bucket_scan_aux_bytes(n))
void launch_exclusive_scan_u32(uint32_t* data, uint32_t n, uint32_t* aux, cudaStream_t s, uint64_t* launches) {
    exclusive_scan(data, n, aux, s, launches);
}

size_t bucket_scan_aux_bytes(uint32_t ncell) {
    uint64_t n = (uint64_t)ncell + 1, tot = 4;
    while (n > 1) { n = (n + kScanTile - 1)/kScanTile; tot += n + 1; }
    return (size_t)tot*sizeof(uint32_t);
}

// Cell tables beyond 2^24 cells (64 MB) take the rank form of the two passes (XM_INDEX_RANK=0 / 1 forces the cursor /
// the rank form).
bool bucket_index_uses_rank(uint32_t ncell) {
    if (const char* e = getenv("XM_INDEX_RANK")) return e[0] == '1';
    return ncell > (1u << 24);
}

// words of the `cursor` scratch: the cursor (one per cell) or the arrival ranks (one per row)
size_t bucket_index_scratch_words(uint32_t n, uint32_t ncell) {
    return bucket_index_uses_rank(ncell) ? (size_t)n : (size_t)ncell + 1;
}

// rec_out: n records; rec_tmp: n records (only touched when order_cells); cstart: ncell+1 words;
// cursor: bucket_index_scratch_words(n, ncell) words; aux: bucket_scan_aux_bytes(ncell). order_cells = false leaves every cell in
// arrival order (callers that do not depend on the order inside a cell skip a full pass);
// *overflow (device word) is incremented when a cell was too large to order.
int32_t launch_bucket_index(const double* ra, const double* dec, uint32_t n, const GridView& g,
                            uint32_t ncell, bool order_cells, Rec* rec_tmp, Rec* rec_out,
                            uint32_t* cstart, uint32_t* cursor, uint32_t* aux, uint32_t* overflow,
                            cudaStream_t s, Error& err, uint64_t* launches) {
    XM_CUDA_TRY(cudaMemsetAsync(cstart, 0, ((size_t)ncell + 1)*sizeof(uint32_t), s));
    const uint32_t blocks = (n + kThreads*kRows - 1)/(kThreads*kRows);
    if (bucket_index_uses_rank(ncell)) {
        uint32_t* rank = cursor;                 // n words (see bucket_index_scratch_words)
        count_rank_kernel<<<blocks, kThreads, 0, s>>>(ra, dec, n, g, cstart, rank);
        *launches += 1;
        exclusive_scan(cstart, ncell + 1, aux, s, launches);
        place_rank_kernel<<<blocks, kThreads, 0, s>>>(ra, dec, n, g, cstart, rank, order_cells ? rec_tmp : rec_out);
        *launches += 1;
    } else {
        count_kernel<<<blocks, kThreads, 0, s>>>(ra, dec, n, g, cstart);
        *launches += 1;
        exclusive_scan(cstart, ncell + 1, aux, s, launches);
        XM_CUDA_TRY(cudaMemcpyAsync(cursor, cstart, (size_t)ncell*sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        place_kernel<<<blocks, kThreads, 0, s>>>(ra, dec, n, g, cursor, order_cells ? rec_tmp : rec_out);
        *launches += 1;
    }
    if (order_cells) {
        const uint64_t threads = (uint64_t)ncell*32;
        order_cells_kernel<<<(unsigned)((threads + kThreads - 1)/kThreads), kThreads, 0, s>>>(rec_tmp, cstart, ncell,
                                                                                            rec_out, overflow);
        *launches += 1;
    }
    XM_CUDA_TRY(cudaGetLastError());
    return XM_OK;
}

}  // namespace xm
