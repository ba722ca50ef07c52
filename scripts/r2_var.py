"""One line of stage times for the C2 workload under the current environment (variant sweeps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vif_b200 import qxmatch, catalogs as G
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
a1, b1 = G.c2(1, n, xp="torch", device="cuda"); a2, b2 = G.c2(2, n, xp="torch", device="cuda")
acc = {}
for k in range(reps):
    r = qxmatch(a1, b1, a2, b2)
    if k >= 3:
        for key in ("ms_total", "ms_bbox", "ms_index", "ms_search", "ms_filter_fwd", "ms_filter_rev"):
            acc.setdefault(key, []).append(r.stats[key])
print(" ".join("%s=%.3f" % (k, sorted(v)[len(v)//2]) for k, v in acc.items()),
      "refined=%d/%d" % (r.stats["rows_refined_fwd"], r.stats["rows_refined_rev"]))
