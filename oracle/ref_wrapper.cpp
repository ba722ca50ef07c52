// TEST INFRASTRUCTURE ONLY -- never linked or imported by the product path.
//
// Thin C entry points around the UNMODIFIED reference implementation
// (vif::astro::qxmatch, /root/reference/include/vif/astro/qxmatch.hpp:133-616).
// The reference headers are used in place (-I/root/reference/include); no reference
// source is copied into this repository. Built by oracle/Makefile into
// oracle/_ref/libqxmatch_ref.so (git-ignored, travels to the GPU box as a binary).
//
// Uses: pinning oracle/qxmatch_oracle.c, minting tests/golden/*.npz, and the CPU
// baseline arm of bench.py (kind = "reference").
#include <vif.hpp>
#include <vif/astro/qxmatch.hpp>

#include <cstdint>
#include <cstring>

using namespace vif;
using namespace vif::astro;

extern "C" {

// Runs the reference qxmatch on raw arrays. Outputs are caller-allocated:
// id,d: nth*n1 (plane-major, qxmatch.hpp:157), rid,rd: n2 (ignored if no_mirror).
// Returns the reference's nth (after the clamp of qxmatch.hpp:152).
uint64_t ref_qxmatch(const double* ra1, const double* dec1, uint64_t n1,
                     const double* ra2, const double* dec2, uint64_t n2,
                     uint64_t nth, uint64_t thread, int self, int brute_force,
                     int no_mirror, int linear,
                     uint64_t* id, double* d, uint64_t* rid, double* rd) {
    vec1d vra1(n1), vdec1(n1), vra2(n2), vdec2(n2);
    if (n1) { std::memcpy(vra1.data.data(), ra1, n1*sizeof(double));
              std::memcpy(vdec1.data.data(), dec1, n1*sizeof(double)); }
    if (n2) { std::memcpy(vra2.data.data(), ra2, n2*sizeof(double));
              std::memcpy(vdec2.data.data(), dec2, n2*sizeof(double)); }

    qxmatch_params p;
    p.thread = thread; p.nth = nth; p.verbose = false; p.self = self != 0;
    p.brute_force = brute_force != 0; p.no_mirror = no_mirror != 0; p.linear = linear != 0;

    qxmatch_res r = qxmatch(vra1, vdec1, vra2, vdec2, p);

    std::memcpy(id, r.id.data.data(), r.id.data.size()*sizeof(uint64_t));
    std::memcpy(d,  r.d.data.data(),  r.d.data.size()*sizeof(double));
    if (!no_mirror) {
        std::memcpy(rid, r.rid.data.data(), r.rid.data.size()*sizeof(uint64_t));
        std::memcpy(rd,  r.rd.data.data(),  r.rd.data.size()*sizeof(double));
    }
    return r.id.dims[0];
}

// field_area_hull of 4 bounding-box corners, as called at qxmatch.hpp:243-247.
double ref_field_area_bbox(double r0, double r1, double d0, double d1) {
    vec1d hx = {r0, r0, r1, r1};
    vec1d hy = {d0, d1, d0, d1};
    return field_area_hull(hx, hy);
}

// angdist in arcsec from degrees (astro.hpp:33-39).
double ref_angdist(double ra1, double dec1, double ra2, double dec2) {
    return angdist(ra1, dec1, ra2, dec2);
}

// vector angdist (astro.hpp:43-81), angdist_less (astro.hpp:83-103), qdist (qxmatch.hpp:684-717)
void ref_angdist_vec(const double* ra1, const double* dec1, uint64_t n, const double* ra2,
                     const double* dec2, uint64_t n2, double* out) {
    vec1d a1(n), b1(n);
    for (uint64_t i = 0; i < n; ++i) { a1[i] = ra1[i]; b1[i] = dec1[i]; }
    vec1d r;
    if (n2 == 1) r = angdist(a1, b1, ra2[0], dec2[0]);
    else {
        vec1d a2(n), b2(n);
        for (uint64_t i = 0; i < n; ++i) { a2[i] = ra2[i]; b2[i] = dec2[i]; }
        r = angdist(a1, b1, a2, b2);
    }
    for (uint64_t i = 0; i < n; ++i) out[i] = r[i];
}

void ref_angdist_less(const double* ra1, const double* dec1, uint64_t n, double ra2, double dec2,
                      double radius, uint8_t* out) {
    vec1d a1(n), b1(n);
    for (uint64_t i = 0; i < n; ++i) { a1[i] = ra1[i]; b1[i] = dec1[i]; }
    vec1b r = angdist_less(a1, b1, ra2, dec2, radius);
    for (uint64_t i = 0; i < n; ++i) out[i] = r[i] ? 1 : 0;
}

void ref_qdist(const double* ra, const double* dec, uint64_t n, double* out) {
    vec1d a(n), b(n);
    for (uint64_t i = 0; i < n; ++i) { a[i] = ra[i]; b[i] = dec[i]; }
    vec2d r = qdist(a, b);
    for (uint64_t k = 0; k < n*n; ++k) out[k] = r.data[k];
}

// Ring table dump (depth_cache, qxmatch.hpp:44-128). Writes ring `k` of a grid
// (nx, ny, cs) into bx/by (capacity cap); returns the entry count, max_dist in *md.
uint64_t ref_ring(uint64_t nx, uint64_t ny, double cs, uint64_t k,
                  int64_t* bx, int64_t* by, uint64_t cap, double* md) {
    impl::qxmatch_impl::depth_cache dc(nx, ny, cs);
    const auto& dep = dc[k];
    uint64_t n = dep.bx.size();
    for (uint64_t i = 0; i < n && i < cap; ++i) { bx[i] = dep.bx[i]; by[i] = dep.by[i]; }
    *md = dep.max_dist;
    return n;
}

uint64_t ref_hardware_threads() { return std::thread::hardware_concurrency(); }

}

// vif's main() wrapper (core/main.hpp:13-22) expects this symbol; unused in the .so.
int vif_main(int, char*[]) { return 0; }
