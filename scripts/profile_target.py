"""Small driver for ncu captures: N device-resident qxmatch calls on the C2 workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vif_b200 import qxmatch, catalogs as G
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
a1, b1 = G.c2(1, n, xp="torch", device="cuda"); a2, b2 = G.c2(2, n, xp="torch", device="cuda")
for _ in range(reps):
    r = qxmatch(a1, b1, a2, b2)
print(r.stats)
