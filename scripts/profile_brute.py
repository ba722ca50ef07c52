import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vif_b200 import qxmatch, catalogs as G
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
a1, b1 = G.fullsky(1, n, xp="torch", device="cuda"); a2, b2 = G.fullsky(2, n, xp="torch", device="cuda")
for _ in range(2):
    r = qxmatch(a1, b1, a2, b2, brute_force=True)
st = r.stats
print({k: v for k, v in st.items() if k != "grid"})
print("pairs/s fwd %.3e" % (n*n/(st["ms_forward"]*1e-3)))
