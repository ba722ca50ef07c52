"""vif_b200 -- B200-native (sm_100a) replacement for vif's catalog cross-match hot path.

Only what `vif::astro::qxmatch` needs lives here:

  csrc/            hand-written CUDA kernels + the C ABI (include/vif_b200_qxmatch.h)
  lib/             the built shared library (git-ignored, built in-tree by build.py / make)
  qxmatch.py       Python mirror of the reference interface (same keywords, same result fields)
  distributed.py   row-sharded multi-GPU driver (one process per GPU, NCCL through torch.distributed)
  dist.py          angdist / angdist_less / qdist on the GPU (astro.hpp:33-103, qxmatch.hpp:684-717)
  catalogs.py      counter-based synthetic catalog generators of SURVEY.md section 8(d)

There is no CPU implementation in this package: without the CUDA library and a GPU every compute
entry point raises.
"""
from .qxmatch import (QxmatchParams, QxmatchRes, QxmatchError, IdPair, qxmatch, xmatch_clean_best,  # noqa: F401
                      last_stats, release_memory, NPOS)

from .dist import angdist, angdist_less, qdist  # noqa: F401

__version__ = "0.1"
