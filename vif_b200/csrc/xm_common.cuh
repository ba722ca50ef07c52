// Shared declarations of the qxmatch device pipeline (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/vif_b200_qxmatch.h"

namespace xm {

// ---- constants (math/base.hpp:20, qxmatch.hpp:174) -----------------------------------------
// dpi is written with 18 digits in the reference and rounds to the double nearest pi.
#define XM_DPI 3.14159265358979323
#define XM_D2R (XM_DPI/180.0)

// ---- device-side views ---------------------------------------------------------------------
// One source of a catalog sorted by cell key: 32 bytes, fetched with a single LDG.256.
// x,y are radians (raw units when linear), c = cos(y).
struct __align__(32) Rec {
    double   x, y, c;
    uint32_t idx;           // original row of this source
    uint32_t key;           // its cell key (ix*ndec + iy)
};

struct CatView {
    const Rec*   rec;       // sources in cell order (stable: ascending row inside a cell)
    const uint32_t* cstart; // CSR: cell k holds sorted slots [cstart[k], cstart[k+1]); ncell+1 entries
    uint32_t     n;
};

struct GridView {
    double   rra0, rdec0;     // padded origin [deg, or raw units when linear] -- binning units
    double   dra, ddec;       // cell widths in binning units
    double   cell_size;       // [arcsec, or raw units when linear]
    double   ox, oy, wx, wy;  // the same grid in search units (radians, or raw units when linear)
    double   fzx, fzy;        // slack on cell edges in search units: covers the rounding of the
                              // binning expression so that cell lower bounds stay rigorous
    const double* cmin_row;   // per cell row: lower bound of cos(dec) over the row (spherical)
    const double* ccen_row;   // per cell row: cos(dec) of the row centre (filter kernel's stop rule)
    double   cs_rad;          // cell_size in radians
    double   eps_lo;          // filter metric Q vs 4*proxy: 4p in [Q(1-eps_lo), Q(1+eps_hi)]
    double   eps_hi;
    double   stop_delta;      // slack on the stop-rule radius [rad]
    uint32_t nra, ndec;
    uint32_t max_depth;       // max(nra, ndec): depth_cache::size(), qxmatch.hpp:85-87
    int32_t  linear;
    int32_t  sin_model;       // 1: glibc __sin_fma Taylor branch, 2: non-FMA build of the same
    uint32_t self_off;        // partitioned self match: source row r is candidate row r + self_off
    unsigned long long* off_grid;   // index build: counter of rows that do not fall on the grid (may be null)
};

struct SearchCounters {       // device counters, zeroed per call
    unsigned long long evals_fwd, evals_rev, ambiguous, refine_fwd, refine_rev;
    unsigned int index_overflow;  // cells too large for the bucket index's in-cell ordering
    unsigned int pad;
    unsigned long long off_grid1, off_grid2;   // rows of catalog 1 / 2 that are not finite or fall outside the grid
                                               // (only possible when the caller promised the bounding boxes)
};

// ---- error plumbing ------------------------------------------------------------------------
struct Error {
    int32_t code = XM_OK;
    char    msg[512] = {0};
};

#define XM_CUDA_TRY(expr)                                                                        \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            return xm::fail_cuda(err, _e, #expr, __FILE__, __LINE__);                            \
        }                                                                                        \
    } while (0)

int32_t fail_cuda(Error& err, cudaError_t e, const char* what, const char* file, int line);
int32_t fail(Error& err, int32_t code, const char* fmt, ...);

// ---- host geometry (xm_geometry.cpp) ------------------------------------------------------
void grid_from_bbox(const double* bb1, const double* bb2, uint64_t n2, uint64_t nth, bool linear,
                    xm_grid* g);
int  calibrate_sin_model();   // probes the host libm: 1 = glibc 2.39 FMA build, 2 = plain build, 0 = not recognised
double libm_host_eval(int which, int model, double x);   // host side of the restatement (0 = sin, 1 = cos)

// ---- launchers -----------------------------------------------------------------------------
// xm_index.cu
int32_t launch_bbox(const double* ra, const double* dec, uint64_t n, unsigned long long* out5,
                    cudaStream_t s, Error& err, uint64_t* launches);
// distance helpers (xm_dist.cu)
int32_t launch_angdist(const double* ra1, const double* dec1, const double* ra2, const double* dec2, uint64_t n,
                       bool scalar2, int model, double* out, cudaStream_t s, Error& err);
int32_t launch_angdist_less(const double* ra1, const double* dec1, uint64_t n, double tra2, double tdec2,
                            double radius, int model, uint8_t* out, cudaStream_t s, Error& err);
size_t  qdist_scratch_bytes(uint64_t n);
int32_t launch_qdist(const double* ra, const double* dec, uint64_t n, int model, void* scratch, double* out,
                     cudaStream_t s, Error& err);

// xm_band.cu: order-preserving declination-band partition (multi-GPU exchange step)
size_t  band_scratch_bytes(uint64_t n, uint64_t cap);
int32_t launch_band_partition(const double* ra, const double* dec, uint32_t n, uint64_t row0, const double* edges,
                              int nb, double halo, int only, double* out, uint64_t cap, unsigned long long* counts,
                              void* scratch, uint64_t* total, cudaStream_t s, Error& err, uint64_t* launches);
int32_t launch_band_finish(long long* id_io, const double* d, const long long* bits_src, const long long* bits_cand,
                           uint32_t n, double hlim, unsigned long long* bad, cudaStream_t s, Error& err);
int32_t launch_band_home(const double* got, const long long* sent, uint32_t m, uint64_t row0, long long* id_out,
                         double* d_out, cudaStream_t s, Error& err);
int32_t launch_libm_eval(int which, int model, const double* x, uint64_t n, double* out, cudaStream_t s, Error& err);
// xm_peak.cu
int32_t launch_fp64_peak(int reps, double* tflops, cudaStream_t s, Error& err);

// pageable host buffers (xm_staging.cu)
struct StageJob { void* dev; void* host; size_t bytes; };
bool    host_pointer_is_pageable(const void* p);
int32_t stage_h2d(int device, const StageJob* jobs, int njobs, cudaStream_t s, Error& err);
int32_t stage_d2h(int device, const StageJob* jobs, int njobs, cudaStream_t s, Error& err);
void    stage_shutdown();

int32_t launch_merge_p2p(const void* const* exch, void* const* out, uint64_t n, int rank, int world,
                         double unmatched, cudaStream_t s, Error& err);
int32_t launch_merge_step(int which, unsigned long long* rid, double* rd, const double* ka, const double* kb,
                          double* key_out, uint64_t n, unsigned long long off, double unmatched,
                          cudaStream_t s, Error& err);
int32_t launch_bbox_neg(const double* ra, const double* dec, uint64_t n, double* out5, cudaStream_t s, Error& err,
                        uint64_t* launches);
int32_t launch_keys(const double* ra, const double* dec, uint32_t n, const GridView& g,
                    uint32_t* keys, uint32_t* vals, cudaStream_t s, Error& err, uint64_t* launches);
int32_t launch_cells_and_gather(const uint32_t* skeys, const uint32_t* perm, const double* ra,
                                const double* dec, uint32_t n, uint32_t ncell, bool linear,
                                Rec* rec, uint32_t* cstart, cudaStream_t s, Error& err,
                                uint64_t* launches, int model, const GridView& g);
// xm_bucket.cu: counting-sort index (count -> scan -> place -> order cells); fast path
size_t  bucket_scan_aux_bytes(uint32_t ncell);
bool    bucket_index_uses_rank(uint32_t ncell);
size_t  bucket_index_scratch_words(uint32_t n, uint32_t ncell);
void    launch_exclusive_scan_u32(uint32_t* data, uint32_t n, uint32_t* aux, cudaStream_t s, uint64_t* launches);
int32_t launch_bucket_index(const double* ra, const double* dec, uint32_t n, const GridView& g,
                            uint32_t ncell, bool order_cells, Rec* rec_tmp, Rec* rec_out,
                            uint32_t* cstart, uint32_t* cursor, uint32_t* aux, uint32_t* overflow,
                            cudaStream_t s, Error& err, uint64_t* launches);
size_t  clean_best_scratch_bytes(uint32_t n1);
int32_t launch_clean_best(const uint64_t* id, const uint64_t* rid, uint32_t n1, uint64_t n2, uint64_t* id1,
                          uint64_t* id2, uint64_t* lost, uint64_t* counts, void* scratch, cudaStream_t s,
                          Error& err, uint64_t* launches);
// xm_sort.cu: stable LSD radix sort, result ends in (keys, vals); tmp_* same size; scratch sized
// by sort_scratch_bytes(n)
size_t  sort_scratch_bytes(uint64_t n);
int32_t launch_sort_pairs(uint32_t* keys, uint32_t* vals, uint32_t* tmp_keys, uint32_t* tmp_vals,
                          uint32_t n, int key_bits, void* scratch, cudaStream_t s, Error& err,
                          uint64_t* launches);
// xm_search_exact.cu: reference-order arithmetic (compiled with -fmad=false)
// self: 0 = off, 1 = both views are the same index (skip equal slots), 2 = skip equal original rows
int32_t launch_ring_search_exact(const GridView& g, const CatView& src, const CatView& cand,
                                 int self, uint32_t nth_eff, uint64_t plane_stride,
                                 uint64_t* id_out, double* d_out, const uint32_t* rows,
                                 uint32_t nrows, const unsigned long long* nrows_dev, bool reverse,
                                 bool cells_unordered, SearchCounters* ctr, cudaStream_t s,
                                 Error& err, uint64_t* launches);
int32_t launch_row_tables(const GridView& g, double* cmin_row, double* ccen_row, cudaStream_t s,
                          Error& err, uint64_t* launches);
// xm_search_fast.cu: nearest-neighbour filter kernel (nth = 1, spherical, rings 0-1); rows it cannot
// decide rigorously are appended to refine_list (count in *refine_count) for the exact kernel
bool    fast_path_applicable(const GridView& g);
size_t  filter_result_bytes(uint64_t n);     // scratch for the filter's sector-sized result records
int32_t launch_unpack_results(const void* res, uint32_t n, uint64_t* id_out, double* d_out, cudaStream_t s,
                              Error& err, uint64_t* launches);
size_t  filter_plan_bytes(uint64_t n);       // scratch for the per-tile staging plans
// xm_search_nn32.cu: the same filter with an FP32 metric in a per-tile frame (catalogs of ordinary
// density: a tile's 3-column neighbourhood must fit the shared-memory stage)
int     nn32_tile_sources(const GridView& g, uint64_t n_src, uint64_t n_cand);   // 0: not applicable
size_t  nn32_plan_bytes(uint64_t n_src, uint32_t nra, int tile);
int32_t launch_ring_nn32(const GridView& g, const CatView& src, const CatView& cand, bool self, int tile,
                         uint64_t* id_out, double* d_out, void* result_scratch, void* plan_scratch,
                         uint32_t* refine_list, unsigned long long* refine_count, bool reverse,
                         SearchCounters* ctr, cudaStream_t s, cudaEvent_t filter_done, Error& err,
                         uint64_t* launches, cudaEvent_t k0 = nullptr, cudaEvent_t k1 = nullptr,
                         cudaEvent_t after = nullptr, bool defer_unpack = false);
int     nn32_persistent_tile();          // tile size served by the persistent form of the kernel (0: switched off)
int32_t launch_ring_nn_filter(const GridView& g, const CatView& src, const CatView& cand, bool self,
                              uint64_t* id_out, double* d_out, void* result_scratch, void* plan_scratch,
                              uint32_t* refine_list, unsigned long long* refine_count, bool reverse,
                              SearchCounters* ctr, cudaStream_t s, cudaEvent_t filter_done, Error& err,
                              uint64_t* launches);
int32_t launch_brute_exact(const GridView& g, const double* x1, const double* y1, const double* c1,
                           uint32_t n1, const double* x2, const double* y2, const double* c2,
                           uint32_t n2, bool self, uint32_t nth_eff, uint64_t plane_stride,
                           uint64_t* id_out, double* d_out, const uint32_t* rows, uint32_t nrows,
                           bool reverse, SearchCounters* ctr, cudaStream_t s, Error& err,
                           uint64_t* launches);
// xm_search_topk.cu: nth = 2..8 filter kernel (multiset rule of the ring table), forward only
bool    topk_filter_applicable(uint32_t nth_eff);
int32_t launch_ring_topk_filter(const GridView& g, const CatView& src, const CatView& cand, bool self,
                                uint32_t nth_eff, uint64_t plane_stride, uint64_t* id_out, double* d_out,
                                uint32_t* refine_list, unsigned long long* refine_count, SearchCounters* ctr,
                                cudaStream_t s, Error& err, uint64_t* launches);
// xm_brute_fast.cu: dot-product filter + exact refine for brute_force (spherical, nth <= 8)
size_t  brute_uvec_bytes(uint64_t n);
bool    brute_fast_applicable(bool linear, uint32_t nth);
int32_t launch_unit_vectors(const double* x, const double* y, const double* c, uint32_t n, void* out,
                            cudaStream_t s, Error& err, uint64_t* launches);
int32_t launch_brute_fast(int sin_model, const void* u1, uint32_t n1, const void* u2, uint32_t n2,
                          const double* x1, const double* y1, const double* c1, const double* x2,
                          const double* y2, const double* c2, bool self, uint32_t nth,
                          uint64_t plane_stride, uint64_t* id_out, double* d_out, uint32_t* refine_list,
                          double* refine_thr, unsigned long long* refine_count, SearchCounters* ctr,
                          cudaStream_t s, Error& err, uint64_t* launches);
// xm_sphere.cu: exact nth-nearest search with a wrap- and pole-aware grid (brute-force semantics)
bool    sphere_grid(const double* bb, uint64_t n_cand, uint32_t nth, GridView* g, int* wrap);
bool    sphere_applicable(bool linear, uint32_t nth);
int32_t launch_sphere_rows(const GridView& g, double* cmin_row, cudaStream_t s, Error& err, uint64_t* launches);
int32_t launch_sphere_search(const GridView& g, int wrap, const double* cmin_row, const CatView& src,
                             const CatView& cand, bool self, uint32_t nth_eff, uint64_t plane_stride,
                             uint64_t* id_out, double* d_out, uint32_t* refine_list,
                             unsigned long long* refine_count, void* result_records, bool reverse,
                             SearchCounters* ctr, cudaStream_t s, cudaStream_t side, cudaEvent_t ev_fork,
                             cudaEvent_t ev_join, Error& err, uint64_t* launches);
int32_t launch_prep_unsorted(const double* ra, const double* dec, uint32_t n, bool linear,
                             double* x, double* y, double* c, cudaStream_t s, Error& err,
                             uint64_t* launches, int model);
int32_t launch_fill_outputs(uint64_t* id, double* d, uint64_t n, double dval, cudaStream_t s,
                            Error& err, uint64_t* launches);

}  // namespace xm
