// Exercises the C++ drop-in exactly the way qxmatch2 / catalog_merge use the reference
// (tools/qxmatch2/qxmatch2.cpp:128-148, catalog_merge.hpp:194-195): vif vectors in, qxmatch_res
// out. Reads catalogs from a raw float64 file, writes id/d/rid/rd to another; the pytest driver
// (tests/test_gpu_dropin.py) compares the output with the oracle.
//   dropin_test <in.bin> <out.bin> <n1> <n2> <nth> <flags: s=self b=brute m=no_mirror l=linear v=view>
#include <vif.hpp>
#include <vif/astro/qxmatch.hpp>      // resolved to include/vif_b200/vif/astro/qxmatch.hpp by -I order

#include <cstdio>
#include <cstring>

#ifndef VIF_B200_ASTRO_QXMATCH_HPP
#error "the drop-in header was not picked up: put include/vif_b200 before vif's include directory"
#endif

using namespace vif;
using namespace vif::astro;

int vif_main(int argc, char* argv[]) {
    if (argc < 7) { print("usage: dropin_test in out n1 n2 nth flags"); return 1; }
    const uint_t n1 = std::stoull(argv[3]), n2 = std::stoull(argv[4]);
    const std::string flags = argv[6];
    vec1d ra1(n1), dec1(n1), ra2(n2), dec2(n2);
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    size_t got = 0;
    got += fread(ra1.data.data(), 8, n1, f); got += fread(dec1.data.data(), 8, n1, f);
    got += fread(ra2.data.data(), 8, n2, f); got += fread(dec2.data.data(), 8, n2, f);
    fclose(f);
    if (got != 2*n1 + 2*n2) return 3;

    qxmatch_params p;
    p.nth = std::stoull(argv[5]);
    p.thread = 4;                                  // as catalog_merge does; ignored by the GPU path
    p.brute_force = flags.find('b') != std::string::npos;
    p.no_mirror = flags.find('m') != std::string::npos;
    p.linear = flags.find('l') != std::string::npos;

    qxmatch_res r;
    if (flags.find('s') != std::string::npos) {
        r = qxmatch(ra1, dec1, p);                 // one-catalog overload (self match)
    } else if (flags.find('v') != std::string::npos) {
        // view arguments + float inputs, as in catalog_merge.hpp:194-195
        vec1u sel = indgen<uint_t>(n1);
        vec1f fra2 = ra2, fdec2 = dec2;
        r = qxmatch(ra1[sel], dec1[sel], fra2, fdec2, p);
    } else {
        struct cat_t { vec1d ra, dec; } c1{ra1, dec1}, c2{ra2, dec2};
        r = qxmatch(c1, c2, p);                    // catalog overload (qxmatch2.cpp:130)
    }

    id_pair best;                                  // mutual-best idiom of the callers (needs rid)
    if (!p.no_mirror && flags.find('s') == std::string::npos) best = xmatch_clean_best(r);
    f = fopen(argv[2], "wb");
    if (!f) return 4;
    uint_t hdr[6] = {r.id.dims[0], r.id.dims[1], r.rid.size(), r.rd.size(), best.id1.size(), best.lost.size()};
    fwrite(hdr, 8, 6, f);
    fwrite(r.id.data.data(), 8, r.id.size(), f); fwrite(r.d.data.data(), 8, r.d.size(), f);
    fwrite(r.rid.data.data(), 8, r.rid.size(), f); fwrite(r.rd.data.data(), 8, r.rd.size(), f);
    fwrite(best.id1.data.data(), 8, best.id1.size(), f);
    fclose(f);
    return 0;
}
