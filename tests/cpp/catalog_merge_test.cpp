// The reference's catalog merger (astro/catalog_merge.hpp:122-245) compiled UNMODIFIED against the
// B200 drop-in: catalog_pool::add_catalog / match_catalog call qxmatch(cra[sel], cdec[sel], ra, dec,
// {nth = 2, thread = 4, verbose}) on VIEWS (catalog_merge.hpp:194-195), cache id/d/rid/rd in a
// hash-named column-oriented FITS file (:197-203) and keep the mutual best pairs (:215-232).
// The header is wrapped in #ifndef NO_CFITSIO because of those cache files; cfitsio is absent here, so
// the vif::fits calls it makes are served by include/vif_b200/io/colfits.hpp: NO_CFITSIO is defined for
// the rest of vif and lifted just for this one header.
//
// usage: catalog_merge_test in.bin out.bin n1 n2 cache_prefix
//   in.bin : ra1[n1] dec1[n1] ra2[n2] dec2[n2] (float64). Catalog 1 seeds the pool, catalog 2 is added.
//   out.bin: u64 ngal, n_idm, then idm[n_idm] (pool row of every catalog-2 source), origin[ngal],
//            f64 ra[ngal], dec[ngal], d(0,:)[n2]
// A second pool then re-reads the cache file (xmatch = false) and must arrive at the same merge.
#include <vif.hpp>
#include <vif/astro/qxmatch.hpp>
// The header as shipped does not compile on its own (g++ 13), cfitsio or not: its impl::astro_impl
// helpers name astro::catalog_t unqualified (:358,379,404) and merge_flux calls strarr(), which no vif
// header declares (:456). Both are bridged from OUTSIDE, the header itself stays untouched:
namespace vif {
    namespace astro { struct catalog_t; }
    namespace impl { namespace astro_impl { using vif::astro::catalog_t; } }
    inline vec1s strarr(uint_t n) { return vec1s(n); }
}
#undef NO_CFITSIO
#include <vif/astro/catalog_merge.hpp>
#define NO_CFITSIO

#include <cstdio>

using namespace vif;
using namespace vif::astro;

int vif_main(int argc, char* argv[]) {
    if (argc < 6) { print("usage: catalog_merge_test in.bin out.bin n1 n2 cache_prefix"); return 2; }
    uint_t n1 = 0, n2 = 0;
    from_string(argv[3], n1); from_string(argv[4], n2);
    vec1d ra1(n1), dec1(n1), ra2(n2), dec2(n2);
    FILE* f = std::fopen(argv[1], "rb");
    if (!f) { print("cannot open input"); return 2; }
    bool ok = std::fread(ra1.data.data(), 8, n1, f) == n1 && std::fread(dec1.data.data(), 8, n1, f) == n1 &&
              std::fread(ra2.data.data(), 8, n2, f) == n2 && std::fread(dec2.data.data(), 8, n2, f) == n2;
    std::fclose(f);
    if (!ok) { print("short input"); return 2; }

    vec1u idm_first;
    for (int pass = 0; pass < 2; ++pass) {
        // pass 0 matches and writes the cache file, pass 1 (xmatch = false) must read it back
        catalog_pool pool(pass == 0, argv[5]);
        pool.add_catalog(ra1, dec1, vec1u{}, "first", {"src1"}, {"file1"});
        catalog_t& c2 = pool.add_catalog(ra2, dec2, vec1u{}, "second", {"src2"}, {"file2"});
        if (pass == 0) {
            idm_first = c2.idm;
            FILE* o = std::fopen(argv[2], "wb");
            uint_t hdr[2] = {pool.ngal, c2.idm.size()};
            std::fwrite(hdr, 8, 2, o);
            std::fwrite(c2.idm.data.data(), 8, c2.idm.size(), o);
            std::fwrite(pool.origin.data.data(), 8, pool.origin.size(), o);
            std::fwrite(pool.ra.data.data(), 8, pool.ra.size(), o);
            std::fwrite(pool.dec.data.data(), 8, pool.dec.size(), o);
            vec1d d0 = c2.d(0,_);
            std::fwrite(d0.data.data(), 8, d0.size(), o);
            std::fclose(o);
        } else if (count(c2.idm != idm_first) != 0 || c2.idm.size() != idm_first.size()) {
            print("cache round trip changed the merge");
            return 3;
        }
    }
    return 0;
}
