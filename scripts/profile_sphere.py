import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vif_b200 import qxmatch, catalogs as G
n1, n2 = 50_000_000, 5_000_000
a1, b1 = G.fullsky(1, n1, xp="torch", device="cuda"); a2, b2 = G.fullsky(2, n2, xp="torch", device="cuda")
for _ in range(2):
    r = qxmatch(a1, b1, a2, b2)
print({k: v for k, v in r.stats.items() if k != "grid"})
