// Shadow of <vif/astro/qxmatch.hpp>: with -I<repo>/include/vif_b200 placed before vif's own
// include path, the reference's include line resolves here and picks the B200-backed version.
#include "../../astro/qxmatch.hpp"
