// Stable LSD radix sort of (cell key, row) u32 pairs -- the "bucket fill" of the reference
// (qxmatch.hpp:270-288: ids are pushed in ascending row order, so order inside a cell must be
// stable) rebuilt as an HBM-streaming sort.
//
// Per digit pass (<= 8 bits):
//   1. tile_histogram: one CTA per tile of kTile pairs counts its digits in shared memory and
//      writes them digit-major into table[digit][tile];
//   2. scan_rows: one CTA per digit turns its row into exclusive prefix sums and emits the digit
//      total;
//   3. tile_scatter: each CTA ranks its tile stably (per-warp match_any ranking, warps own
//      consecutive runs of the tile), reorders the tile by digit in shared memory so that equal
//      digits leave as contiguous, coalesced runs, and writes them at table base + local offset.
// Traffic per pass: 4 B (histogram) + 8 B read + 8 B written per pair.
#include "xm_common.cuh"

namespace xm {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads/32;
constexpr int kItems = 16;                  // pairs per thread
constexpr int kTile = kThreads*kItems;      // 4096 pairs per CTA
constexpr int kDigits = 256;

__global__ void __launch_bounds__(kThreads)
tile_histogram(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t mask,
               uint32_t ntiles, uint32_t* __restrict__ table) {
    __shared__ uint32_t hist[kDigits];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x*kTile;
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const uint32_t i = base + k*kThreads + threadIdx.x;
        if (i < n) atomicAdd(&hist[(__ldg(keys + i) >> shift) & mask], 1u);
    }
    __syncthreads();
    if (threadIdx.x <= mask) table[(size_t)threadIdx.x*ntiles + blockIdx.x] = hist[threadIdx.x];
}

// grid = number of digits; row d of table becomes its exclusive scan, totals[d] its sum
__global__ void __launch_bounds__(kThreads)
scan_rows(uint32_t* __restrict__ table, uint32_t ntiles, uint32_t* __restrict__ totals) {
    __shared__ uint32_t warp_sum[kWarps];
    __shared__ uint32_t carry_s;
    uint32_t* row = table + (size_t)blockIdx.x*ntiles;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint32_t b = 0; b < ntiles; b += kThreads) {
        const uint32_t i = b + threadIdx.x;
        const uint32_t v = i < ntiles ? row[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_sum[w] = inc;
        __syncthreads();
        uint32_t woff = 0;
        for (int k = 0; k < w; ++k) woff += warp_sum[k];
        const uint32_t carry = carry_s;
        if (i < ntiles) row[i] = carry + woff + inc - v;
        __syncthreads();
        if (threadIdx.x == kThreads - 1) carry_s = carry + woff + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) totals[blockIdx.x] = carry_s;
}

__global__ void __launch_bounds__(kThreads)
tile_scatter(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t n,
             int shift, uint32_t mask, uint32_t ntiles, const uint32_t* __restrict__ table,
             const uint32_t* __restrict__ totals, uint32_t* __restrict__ keys_out,
             uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t cnt[kWarps][kDigits];    // per-warp digit counts -> exclusive offsets
    __shared__ uint32_t dstart[kDigits];         // first tile-local slot of each digit
    __shared__ uint32_t gbase[kDigits];          // global slot of the tile's first pair per digit
    __shared__ uint32_t skey[kTile];
    __shared__ uint32_t sval[kTile];

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int k = 0; k < kWarps; ++k) cnt[k][threadIdx.x] = 0;
    __syncthreads();

    // warp w owns the consecutive run [w*kItems*32, (w+1)*kItems*32) of the tile
    const uint32_t wbase = blockIdx.x*kTile + w*(kItems*32);
    uint32_t key[kItems], val[kItems];
    uint16_t rank[kItems];
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const uint32_t i = wbase + k*32 + lane;
        key[k] = i < n ? __ldg(keys_in + i) : 0xffffffffu;
        val[k] = i < n ? __ldg(vals_in + i) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const uint32_t i = wbase + k*32 + lane;
        const bool live = i < n;
        // dead lanes get a digit of their own (> mask) so they never join a live group
        const uint32_t dg = live ? ((key[k] >> shift) & mask) : (kDigits + lane);
        const uint32_t peers = __match_any_sync(0xffffffffu, dg);
        const uint32_t before = __popc(peers & ((1u << lane) - 1u));
        uint32_t prior = 0;
        if (live) prior = cnt[w][dg];
        __syncwarp();
        if (live && before == 0) cnt[w][dg] = prior + __popc(peers);
        __syncwarp();
        rank[k] = (uint16_t)(prior + before);
    }
    __syncthreads();

    // per digit: exclusive scan over warps, tile totals, digit starts, global bases
    {
        const int dg = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < kWarps; ++k) { const uint32_t c = cnt[k][dg]; cnt[k][dg] = run; run += c; }
        // exclusive scan of the tile totals over digits (256 values, one per thread)
        uint32_t inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        __shared__ uint32_t wsum[kWarps];
        __shared__ uint32_t wtot[kWarps];
        if (lane == 31) wsum[w] = inc;
        // same scan for the global digit totals -> digit offsets in the output
        const uint32_t tot = (uint32_t)dg <= mask ? totals[dg] : 0u;
        uint32_t tinc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, tinc, o);
            if (lane >= o) tinc += t;
        }
        if (lane == 31) wtot[w] = tinc;
        __syncthreads();
        uint32_t woff = 0, toff = 0;
        for (int k = 0; k < w; ++k) { woff += wsum[k]; toff += wtot[k]; }
        dstart[dg] = woff + inc - run;
        const uint32_t tb = (uint32_t)dg <= mask ? table[(size_t)dg*ntiles + blockIdx.x] : 0u;
        gbase[dg] = toff + tinc - tot + tb;
    }
    __syncthreads();

    // reorder the tile by digit in shared memory
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const uint32_t i = wbase + k*32 + lane;
        if (i < n) {
            const uint32_t dg = (key[k] >> shift) & mask;
            const uint32_t slot = dstart[dg] + cnt[w][dg] + rank[k];
            skey[slot] = key[k];
            sval[slot] = val[k];
        }
    }
    __syncthreads();

    const uint32_t tile_n = min((uint32_t)kTile, n - blockIdx.x*kTile);
#pragma unroll
    for (int k = 0; k < kItems; ++k) {
        const uint32_t slot = k*kThreads + threadIdx.x;
        if (slot < tile_n) {
            const uint32_t kk = skey[slot];
            const uint32_t dg = (kk >> shift) & mask;
            const uint32_t dst = gbase[dg] + (slot - dstart[dg]);
            keys_out[dst] = kk;
            vals_out[dst] = sval[slot];
        }
    }
}

}  // namespace

size_t sort_scratch_bytes(uint64_t n) {
    const uint64_t ntiles = (n + kTile - 1)/kTile;
    return (size_t)(kDigits*ntiles + kDigits)*sizeof(uint32_t);
}

int32_t launch_sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t* tmp_keys, uint32_t* tmp_vals,
                          uint32_t n, int key_bits, void* scratch, cudaStream_t s, Error& err,
                          uint64_t* launches) {
    if (n == 0) return XM_OK;
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 32) key_bits = 32;
    const uint32_t ntiles = (n + kTile - 1)/kTile;
    uint32_t* table = (uint32_t*)scratch;
    uint32_t* totals = table + (size_t)kDigits*ntiles;

    // split the key into the fewest <= 8-bit digits, evenly
    const int npass = (key_bits + 7)/8;
    uint32_t *kin = keys, *vin = vals, *kout = tmp_keys, *vout = tmp_vals;
    int shift = 0;
    for (int p = 0; p < npass; ++p) {
        const int bits = (key_bits - shift + (npass - p) - 1)/(npass - p);
        const uint32_t mask = (1u << bits) - 1u;
        tile_histogram<<<ntiles, kThreads, 0, s>>>(kin, n, shift, mask, ntiles, table);
        scan_rows<<<mask + 1, kThreads, 0, s>>>(table, ntiles, totals);
        tile_scatter<<<ntiles, kThreads, 0, s>>>(kin, vin, n, shift, mask, ntiles, table, totals,
                                                 kout, vout);
        *launches += 3;
        uint32_t* t;
        t = kin; kin = kout; kout = t;
        t = vin; vin = vout; vout = t;
        shift += bits;
    }
    XM_CUDA_TRY(cudaGetLastError());
    if (kin != keys) {   // odd number of passes: result sits in the tmp buffers
        XM_CUDA_TRY(cudaMemcpyAsync(keys, kin, (size_t)n*sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        XM_CUDA_TRY(cudaMemcpyAsync(vals, vin, (size_t)n*sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    }
    return XM_OK;
}

}  // namespace xm
