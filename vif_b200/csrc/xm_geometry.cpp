// Host-side scalar geometry of the cell-binned path.
//
// These few scalars decide which cell every source falls in, so they are computed on the host
// with the host libm, in the reference's operation order (qxmatch.hpp:232-267, astro.hpp:524-552,
// convex_hull.hpp:108-150): the device kernels then reproduce the reference's cell keys bit for
// bit with nothing but IEEE subtract/divide/floor.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <vector>

#include "xm_common.cuh"
#include "xm_libm_glibc.cuh"

namespace xm {

namespace {

const double kPi = XM_DPI;

struct Pt { double x, y; };

// great-circle separation of two points given in radians (astro.hpp:25-29)
double sep_rad(const Pt& a, const Pt& b) {
    const double sra = std::sin(0.5*(b.x - a.x));
    const double sde = std::sin(0.5*(b.y - a.y));
    return 2.0*std::asin(std::sqrt(sde*sde + sra*sra*std::cos(b.y)*std::cos(a.y)));
}

double turn(const Pt& o, const Pt& a, const Pt& b) {
    return (a.x - o.x)*(b.y - o.y) - (a.y - o.y)*(b.x - o.x);
}

// Area [deg^2] the reference assigns to catalog 2's bounding box: the four corners go through a
// monotone-chain hull (closed, counter-clockwise; collinear and repeated points dropped with
// "<= 0"), the hull is fanned from its first vertex and every triangle is measured with Heron's
// formula on great-circle edge lengths.
double bbox_hull_area(double r0, double r1, double d0, double d1) {
    std::array<Pt, 4> pts = {{{r0, d0}, {r0, d1}, {r1, d0}, {r1, d1}}};
    std::sort(pts.begin(), pts.end(), [](const Pt& a, const Pt& b) {
        return a.x < b.x || (a.x == b.x && a.y < b.y);
    });
    std::vector<Pt> h;
    h.reserve(8);
    for (const Pt& p : pts) {
        while (h.size() >= 2 && turn(h[h.size()-2], h[h.size()-1], p) <= 0) h.pop_back();
        h.push_back(p);
    }
    const size_t lower = h.size() + 1;
    for (int i = 2; i >= 0; --i) {
        const Pt& p = pts[i];
        while (h.size() >= lower && turn(h[h.size()-2], h[h.size()-1], p) <= 0) h.pop_back();
        h.push_back(p);
    }
    const double d2r = kPi/180.0;
    for (Pt& p : h) { p.x *= d2r; p.y *= d2r; }
    double area = 0;
    for (size_t i = 2; i < h.size(); ++i) {
        const double e1 = sep_rad(h[0], h[i-1]);
        const double e2 = sep_rad(h[i-1], h[i]);
        const double e3 = sep_rad(h[i], h[0]);
        const double s = 0.5*(e1 + e2 + e3);
        area += std::sqrt(s*(s - e1)*(s - e2)*(s - e3));
    }
    area /= d2r*d2r;
    return area;
}

}  // namespace

void grid_from_bbox(const double* bb1, const double* bb2, uint64_t n2, uint64_t nth, bool linear,
                    xm_grid* g) {
    double ra_lo = std::min(bb1[0], bb2[0]), ra_hi = std::max(bb1[1], bb2[1]);
    double de_lo = std::min(bb1[2], bb2[2]), de_hi = std::max(bb1[3], bb2[3]);

    // qxmatch.hpp:241-242: about 10 catalog-2 sources per cell per requested neighbour
    const double overgrowth = 10.0;
    const uint64_t nc2 = (uint64_t)std::ceil(0.5*std::sqrt(kPi*(double)n2/(double)nth/overgrowth));

    double area = 0.0, cs, wra, wdec;
    if (linear) {
        cs = std::sqrt((bb2[1] - bb2[0])*(bb2[3] - bb2[2]));
        wra = cs; wdec = cs;
    } else {
        area = bbox_hull_area(bb2[0], bb2[1], bb2[2], bb2[3]);
        cs = std::max(3600.0*std::sqrt(area)/(double)nc2, 1.0);
        const double mid_dec = (0.0 + bb2[2] + bb2[3])/2.0;       // mean() of {min, max}
        wra = cs/std::fabs(std::cos(mid_dec*kPi/180.0))/3600.0;
        wdec = cs/3600.0;
    }
    ra_lo -= wra; ra_hi += wra; de_lo -= wdec; de_hi += wdec;   // one cell of padding

    g->rra0 = ra_lo; g->rra1 = ra_hi; g->rdec0 = de_lo; g->rdec1 = de_hi;
    g->cell_size = cs; g->dra = wra; g->ddec = wdec;
    const double fx = (ra_hi - ra_lo)/wra, fy = (de_hi - de_lo)/wdec;
    // double -> uint_t truncation; saturate instead of invoking UB on absurd grids
    g->nra  = (fx >= 0 && fx < 1.8e19) ? (uint64_t)fx : UINT64_MAX;
    g->ndec = (fy >= 0 && fy < 1.8e19) ? (uint64_t)fy : UINT64_MAX;
    g->nc2 = nc2; g->area = area;
}

// Which build of glibc 2.39's sin/cos does this host's libm follow? 1: the -mfma build (contracted
// a*b+c), 2: the plain build, 0: neither (another libm or glibc version) -- then the device uses the
// CUDA libm and near ties are only counted (xm_stats.rows_ambiguous). Every probe must agree: the
// arguments cover the Taylor branch of sin, both table branches and the pi/2 - |x| branches of both.
int calibrate_sin_model() {
    static const double probes[] = {
        0x1.e4d7129b19786p-4, -0x1.9a74be59977e6p-5, 0x1.73d97b094ef53p-4, 0x1.d1a8d18d90886p-4,
        -0x1.f649c97d63135p-4, 0x1.9e46618447c9dp-4, -0x1.e72bb4939298ap-4, 0x1.c677bc6a3ec59p-4,
        0x1.b1c4f64361285p-4, 0x1.dbe65e5e43c19p-4, -0x1.b8a4933f18125p-4, -0x1.d2cddbf963a44p-4,
        1e-9, 3.3e-5, -0.0021, 0.0871557427, 0.13, -0.2617993878, 0.3141592, 0.5, -0.61, 0.7853981634, 0.85,
        0.86, -0.9, 1.0, 1.0471975512, -1.2, 1.3962634016, 1.5, 1.5707963267948966, -1.5707, 1.7, 2.0, -2.3, 2.42};
    bool ok[3] = {false, true, true};
    unsigned seed = 12345u;
    for (int it = 0; it < 4096; ++it) {
        double x;
        if (it < (int)(sizeof(probes)/sizeof(probes[0]))) x = probes[it];
        else {                                   // pseudo-random arguments in (-2.42, 2.42), denser near 0
            seed = seed*1664525u + 1013904223u;
            const double u = (double)(seed >> 8)/16777216.0;
            seed = seed*1664525u + 1013904223u;
            const double v = (double)(seed >> 8)/16777216.0;
            x = (2.0*u - 1.0)*2.42*(it % 3 == 0 ? v*v : 1.0);
        }
        volatile double vx = x;                  // keep the libm calls
        const double rs = std::sin(vx), rc = std::cos(vx);
        for (int m = 1; m <= 2; ++m) {
            double es = 0.0, ec = 0.0;
            if (!glibc::sin_small(m, x, &es) || !glibc::cos_small(m, x, &ec) || es != rs || ec != rc) ok[m] = false;
        }
    }
    if (ok[1]) return 1;
    if (ok[2]) return 2;
    return 0;
}

double libm_host_eval(int which, int model, double x) {
    double r = NAN;
    if (which == 0) { if (!glibc::sin_small(model, x, &r)) r = NAN; }
    else if (!glibc::cos_small(model, x, &r)) r = NAN;
    return r;
}

}  // namespace xm
