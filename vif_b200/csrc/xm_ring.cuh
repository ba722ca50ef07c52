// Ring geometry of the reference's depth_cache (qxmatch.hpp:44-128) in closed form.
//
// The reference grows a table: ring 0 = {(0,0)}; ring k >= 1 (its "d" = k+1 =: D) takes, in one
// quadrant, the not-yet-visited cells (x,y), x,y in [0,D], whose nearest corner lies within
// cs*(D+0.5) of the origin cell's centre, scanning x outer / y inner ascending, then mirrors the
// quadrant list to (-x,y), (-x,-y), (x,-y) -- so cells on an axis appear twice.
//
// In units of half cells the nearest-corner distance^2 of cell (x,y) is X^2 + Y^2 with
// X = (x == 0 ? 1 : 2x-1), likewise Y: both odd, so the sum is 2 (mod 8), whereas the ring
// radii^2 are (2D+1)^2 = 1 (mod 8) -- the comparison never ties, so the table does not depend on cs at all and
// ring k >= 2 is exactly the annulus (2D-1)^2 < X^2 + Y^2 <= (2D+1)^2; ring 1 is the 3x3 quadrant
// block minus the origin (ring 0 only ever holds the origin cell, whatever its radius).
// Nothing needs to be stored: the kernels enumerate rings on the fly, in the reference's order.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define XM_HD __host__ __device__ __forceinline__
#else
#define XM_HD inline
#endif

namespace xm {

XM_HD int64_t isqrt_floor(int64_t t) {
    int64_t r = (int64_t)sqrt((double)t);
    while (r*r > t) --r;
    while ((r + 1)*(r + 1) <= t) ++r;
    return r;
}

// y-interval [ylo, yhi] of ring k >= 1 at quadrant column x (0 <= x <= k+1). false when empty.
XM_HD bool ring_yrange(uint32_t k, uint32_t x, uint32_t& ylo, uint32_t& yhi) {
    if (k == 1) {
        if (x > 2) return false;
        ylo = (x == 0) ? 1u : 0u;
        yhi = 2u;
        return true;
    }
    const int64_t D = (int64_t)k + 1;
    const int64_t X = (x == 0) ? 1 : 2*(int64_t)x - 1;
    const int64_t t1 = (2*D - 1)*(2*D - 1) - X*X;     // need Y^2 >  t1
    const int64_t t2 = (2*D + 1)*(2*D + 1) - X*X;     // need Y^2 <= t2
    if (t2 < 1) return false;
    int64_t ymax = isqrt_floor(t2);
    if ((ymax & 1) == 0) --ymax;                      // largest odd Y, >= 1
    int64_t ymin = 1;
    if (t1 >= 1) {
        ymin = isqrt_floor(t1) + 1;
        if ((ymin & 1) == 0) ++ymin;                  // smallest odd Y with Y^2 > t1
    }
    if (ymin > ymax) return false;
    ylo = (ymin == 1) ? 0u : (uint32_t)((ymin + 1)/2); // Y = 1 stands for both y = 0 and y = 1
    yhi = (uint32_t)((ymax + 1)/2);
    return true;
}

// Calls visit(bx, by) for every entry of ring k of an (nx, ny) grid in the reference's order.
// dedupe = true skips the second occurrence of axis cells (only valid where a repeated visit
// provably cannot change the outcome, i.e. nearest-neighbour searches with strict '<').
#ifdef __CUDACC__
#pragma nv_exec_check_disable
#endif
template <class F>
XM_HD void for_each_ring_entry(uint32_t k, uint32_t nx, uint32_t ny, bool dedupe, F&& visit) {
    if (k == 0) { visit(0, 0); return; }
    const uint32_t D = k + 1;
    const uint32_t xm = D < nx - 1 ? D : nx - 1;      // x < min(D+1, nx)
    const uint32_t ym = D < ny - 1 ? D : ny - 1;
    for (int quad = 0; quad < 4; ++quad) {
        const int sx = (quad == 1 || quad == 2) ? -1 : 1;
        const int sy = (quad >= 2) ? -1 : 1;
        for (uint32_t x = 0; x <= xm; ++x) {
            if (dedupe && x == 0 && (quad == 1 || quad == 3)) continue;
            uint32_t ylo, yhi;
            if (!ring_yrange(k, x, ylo, yhi)) continue;
            if (yhi > ym) yhi = ym;
            for (uint32_t y = ylo; y <= yhi; ++y) {
                if (dedupe && y == 0 && (quad == 2 || quad == 3)) continue;
                visit(sx*(int)x, sy*(int)y);
            }
        }
    }
}

}  // namespace xm
