"""Small driver for compute-sanitizer: every code path once, tiny sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vif_b200 import qxmatch, catalogs as G
a1, b1 = G.c1(1, 6000); a2, b2 = G.c1(2, 5000)
qxmatch(a1, b1, a2, b2)                        # bucket index (unordered) + filter + exact refine
qxmatch(a1, b1, a2, b2, nth=3)                 # ordered bucket index + exact kernel
qxmatch(a1, b1, nth=2)                         # self
qxmatch(a1, b1, a2, b2, nth=10)                # MemList
qxmatch(a1[:800], b1[:800], a2[:700], b2[:700], linear=True)
f1 = G.fullsky(3, 3000); f2 = G.fullsky(4, 2500)
qxmatch(*f1, *f2, brute_force=True)            # dot filter + TMA tiles + refine
qxmatch(*f1, *f2, brute_force=True, nth=6)     # exact brute
qxmatch(*f1, *f2, brute_force=True, exact_grid=True, nth=2)   # sphere search
os.environ["XM_INDEX"] = "radix"
qxmatch(a1, b1, a2, b2, nth=2)                 # radix index
print("done")
