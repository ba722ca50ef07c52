// Drop-in replacement for <vif/astro/qxmatch.hpp> backed by the B200 library.
//
// Same namespace, same types, same signatures as the reference header
// (/root/reference/include/vif/astro/qxmatch.hpp): qxmatch_res (:8-17), qxmatch_params (:31-39),
// qxmatch(ra1, dec1, ra2, dec2, params) (:133-136) and its three overloads (:618-634), plus the
// host-side helpers that live in the same file (xmatch_clean_best :641-663, xmatch_check_lost
// :665-669, xmatch_save_lost :672-677, qdist :684-717) so that replacing the file loses nothing.
//
// Usage: put <repo>/include/vif_b200 BEFORE vif's own include directory (-I order) and keep
//     #include <vif.hpp>
//     #include <vif/astro/qxmatch.hpp>          // now resolves to vif_b200/vif/astro/qxmatch.hpp
// or include "vif_b200/astro/qxmatch.hpp" explicitly; link with -lvif_b200_qxmatch.
// The reference's include guard is claimed here, so a later include of the original is a no-op.
//
// What changes underneath: the search itself runs on the GPU through the C ABI of
// include/vif_b200_qxmatch.h (xm_qxmatch_f64). Input checks and error behaviour stay those of the
// reference: vif_check prints the message and exits (core/error.hpp:330-335). `thread` is accepted
// and ignored. There is no CPU fallback: without the library / a GPU the call fails loudly.
#ifndef VIF_ASTRO_QXMATCH_HPP
#define VIF_ASTRO_QXMATCH_HPP
#define VIF_B200_ASTRO_QXMATCH_HPP

#include "vif/astro/astro.hpp"
#include "vif_b200_qxmatch.h"
// without cfitsio the reference has no vif::fits at all: the column-oriented tables of this path
// (RA/DEC in, ID/D/RID/RD out, catalog_merge's cache files) are read and written by the shim
#include "../io/colfits.hpp"

#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

namespace vif {
namespace astro {
    // field-for-field the reference's result type (qxmatch.hpp:8-17)
    struct qxmatch_res {
        vec2u id;
        vec2d d;
        vec1u rid;
        vec1d rd;

        MEMBERS1(id, d, rid, rd);
        MEMBERS2("qxmatch_res", MAKE_MEMBER(id), MAKE_MEMBER(d), MAKE_MEMBER(rid), MAKE_MEMBER(rd));
    };

    // same on-disk layout as the reference: one column-oriented table ID, D, RID, RD (qxmatch.hpp:19-29);
    // through cfitsio where the reference was built with it, through io/colfits.hpp otherwise
    inline void qxmatch_save(const std::string& file, const qxmatch_res& r) {
        fits::write_table(file, ftable(r.id, r.d, r.rid, r.rd));
    }

    inline qxmatch_res qxmatch_restore(const std::string& file) {
        qxmatch_res r;
        fits::read_table(file, ftable(r.id, r.d, r.rid, r.rd));
        return r;
    }

    // field-for-field the reference's keyword set (qxmatch.hpp:31-39)
    struct qxmatch_params {
        uint_t thread = 1u;       // accepted, ignored by the GPU path
        uint_t nth = 1u;
        bool verbose = false;
        bool self = false;
        bool brute_force = false;
        bool no_mirror = false;
        bool linear = false;
    };
}

namespace impl {
    namespace qxmatch_b200 {
        static_assert(sizeof(uint_t) == sizeof(std::uint64_t), "uint_t must be 64 bits wide");

        // any vec<1,T> (owning, view, any arithmetic T) -> contiguous double
        template<typename V>
        std::vector<double> to_f64(const V& v) {
            std::vector<double> out(v.size());
            for (uint_t i = 0; i < v.size(); ++i) out[i] = static_cast<double>(v.safe[i]);
            return out;
        }

        inline void check_rc(std::int32_t rc, const char* msg) {
            vif_check(rc == XM_OK, msg);
        }
    }
}

namespace astro {
    template<typename TypeR1, typename TypeD1, typename TypeR2, typename TypeD2>
    qxmatch_res qxmatch(const vec<1,TypeR1>& ra1, const vec<1,TypeD1>& dec1,
        const vec<1,TypeR2>& ra2, const vec<1,TypeD2>& dec2,
        qxmatch_params params = qxmatch_params{}) {

        // same checks, same messages as the reference (qxmatch.hpp:140-149)
        vif_check(ra1.dims == dec1.dims, "first RA and Dec dimensions do not match (",
            ra1.dims, " vs ", dec1.dims, ")");
        vif_check(ra2.dims == dec2.dims, "second RA and Dec dimensions do not match (",
            ra2.dims, " vs ", dec2.dims, ")");
        vif_check(count(!is_finite(ra1) || !is_finite(dec1)) == 0,
            "first RA and Dec coordinates contain invalid values (infinite or NaN)");
        vif_check(count(!is_finite(ra2) || !is_finite(dec2)) == 0,
            "second RA and Dec coordinates contain invalid values (infinite or NaN)");

        const uint_t nth = params.nth < 1u ? 1u : params.nth;
        const uint_t n1 = ra1.size();
        const uint_t n2 = ra2.size();

        // one catalog matched against itself: hand the library the same buffers twice
        const bool same = static_cast<const void*>(&ra1) == static_cast<const void*>(&ra2) &&
                          static_cast<const void*>(&dec1) == static_cast<const void*>(&dec2);
        std::vector<double> a1 = impl::qxmatch_b200::to_f64(ra1);
        std::vector<double> b1 = impl::qxmatch_b200::to_f64(dec1);
        std::vector<double> a2s, b2s;
        if (!same) {
            a2s = impl::qxmatch_b200::to_f64(ra2);
            b2s = impl::qxmatch_b200::to_f64(dec2);
        }
        const std::vector<double>& a2 = same ? a1 : a2s;
        const std::vector<double>& b2 = same ? b1 : b2s;

        qxmatch_res res;
        res.id.resize(nth, n1);
        res.d.resize(nth, n1);
        if (!params.no_mirror) {
            res.rid.resize(n2);
            res.rd.resize(n2);
        }

        xm_params p = xm_params();
        p.struct_size = sizeof(xm_params);
        p.thread = params.thread; p.nth = nth;
        p.verbose = params.verbose; p.self = params.self; p.brute_force = params.brute_force;
        p.no_mirror = params.no_mirror; p.linear = params.linear;
        p.device = -1;

        char err[512] = {0};
        const std::int32_t rc = xm_qxmatch_f64(a1.data(), b1.data(), n1, a2.data(), b2.data(), n2, &p,
            reinterpret_cast<std::uint64_t*>(res.id.data.data()), res.d.data.data(),
            params.no_mirror ? nullptr : reinterpret_cast<std::uint64_t*>(res.rid.data.data()),
            params.no_mirror ? nullptr : res.rd.data.data(), err, sizeof(err));
        impl::qxmatch_b200::check_rc(rc, err);

        return res;
    }

    template<typename C1, typename C2,
        typename enable = typename std::enable_if<!meta::is_vec<C1>::value>::type>
    qxmatch_res qxmatch(const C1& cat1, const C2& cat2, qxmatch_params params = qxmatch_params{}) {
        return qxmatch(cat1.ra, cat1.dec, cat2.ra, cat2.dec, params);
    }

    template<typename TypeR1, typename TypeD1>
    qxmatch_res qxmatch(const vec<1,TypeR1>& ra1, const vec<1,TypeD1>& dec1,
        qxmatch_params params = qxmatch_params{}) {
        params.self = true;
        return qxmatch(ra1, dec1, ra1, dec1, params);
    }

    template<typename C1, typename enable = typename std::enable_if<!meta::is_vec<C1>::value>::type>
    qxmatch_res qxmatch(const C1& cat1, qxmatch_params params = qxmatch_params{}) {
        return qxmatch(cat1.ra, cat1.dec, params);
    }

    // ---- helpers that share the reference header (host side) -----------------------------------
    struct id_pair {
        vec1u id1, id2;
        vec1u lost;
    };

    // mutual best matches: i is kept when its nearest neighbour's nearest neighbour is i
    inline id_pair xmatch_clean_best(const qxmatch_res& r) {
        const uint_t n = r.id.dims[1];
        std::vector<uint_t> keep1, keep2, lost;
        keep1.reserve(n); keep2.reserve(n);
        for (uint_t i = 0; i < n; ++i) {
            const uint_t j = r.id[i];
            if (r.rid[j] == i) { keep1.push_back(i); keep2.push_back(j); }
            else lost.push_back(i);
        }
        id_pair c;
        c.id1.data = keep1; c.id1.dims[0] = keep1.size();
        c.id2.data = keep2; c.id2.dims[0] = keep2.size();
        c.lost.data = lost; c.lost.dims[0] = lost.size();
        return c;
    }

    inline void xmatch_check_lost(const id_pair& p) {
        if (!p.lost.empty()) {
            warning(p.lost.size(), " sources failed to cross match");
        }
    }

    inline void xmatch_save_lost(const id_pair& p, const std::string& save) {
        if (!p.lost.empty()) {
            warning(p.lost.size(), " sources failed to cross match");
            fits::write_table(save, ftable(p.lost));
        }
    }

    struct qdist_params {
        bool verbose = false;
    };

    // lower-triangular matrix of pairwise angular distances [arcsec] (host side, O(n^2))
    template<typename TypeR, typename TypeD>
    vec2d qdist(const vec<1,TypeR>& ra, const vec<1,TypeD>& dec,
        qdist_params params = qdist_params{}) {
        vif_check(ra.dims == dec.dims, "first RA and Dec dimensions do not match (",
            ra.dims, " vs ", dec.dims, ")");
        const uint_t n = ra.size();
        const double d2r = dpi/180.0;
        std::vector<double> x(n), y(n), c(n);
        for (uint_t i = 0; i < n; ++i) {
            x[i] = ra.safe[i]*d2r; y[i] = dec.safe[i]*d2r; c[i] = cos(y[i]);
        }
        vec2d ret(n, n);
        auto pg = progress_start(n*(n-1)/2);
        for (uint_t i = 0; i < n; ++i)
        for (uint_t j = i+1; j < n; ++j) {
            const double sra = sin(0.5*(x[j] - x[i]));
            const double sde = sin(0.5*(y[j] - y[i]));
            ret.safe(j,i) = 3600.0*(180.0/dpi)*2*asin(sqrt(sde*sde + sra*sra*c[j]*c[i]));
            if (params.verbose) progress(pg, 113);
        }
        return ret;
    }
}
}

#endif
