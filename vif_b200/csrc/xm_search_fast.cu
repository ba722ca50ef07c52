// Filter kernels (placeholder translation unit; filled in once the exact path is parity-green).
#include "xm_common.cuh"
namespace xm {}
