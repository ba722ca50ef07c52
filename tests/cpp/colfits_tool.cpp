// Round-trip helper for tests of include/vif_b200/io/colfits.hpp (no GPU involved):
//   colfits_tool in.fits out.fits    reads RA (1-d double), DEC (1-d, converted to double from whatever the
//                                    file holds), optional ID (2-d uint) and N (scalar uint) and writes them back
#include <vif.hpp>
#include <vif/astro/qxmatch.hpp>

using namespace vif;

int vif_main(int argc, char* argv[]) {
    if (argc < 3) return 2;
    struct { vec1d ra, dec; } cat;
    fits::read_table(argv[1], "ra", cat.ra, "dec", cat.dec);
    vec2u id;
    uint_t n = 0;
    bool more = argc > 3;
    if (more) fits::read_table(argv[1], ftable(id, n));
    if (more) fits::write_table(argv[2], ftable(cat.ra, cat.dec, id, n));
    else fits::write_table(argv[2], "ra", cat.ra, "dec", cat.dec);
    print(cat.ra.size(), " ", cat.dec.size(), " ", id.dims[0], "x", id.dims[1], " ", n);
    return 0;
}
