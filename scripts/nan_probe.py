"""Which launch trips over a non-finite row when the bounding boxes are promised? (CUDA_LAUNCH_BLOCKING=1)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vif_b200 import qxmatch, catalogs as G
from vif_b200.qxmatch import QxmatchParams
kw = eval(sys.argv[1]) if len(sys.argv) > 1 else dict(nth=3)
a1, b1 = [torch.from_numpy(x).cuda() for x in G.c2(91, 150_000)]
a2, b2 = [torch.from_numpy(x).cuda() for x in G.c2(92, 130_000)]
bb1 = (float(a1.min()), float(a1.max()), float(b1.min()), float(b1.max()))
bb2 = (float(a2.min()), float(a2.max()), float(b2.min()), float(b2.max()))
bad = a2.clone(); bad[7] = float("nan")
try:
    qxmatch(a1, b1, bad, b2, params=QxmatchParams(bbox1=bb1, bbox2=bb2, **kw))
    print(kw, "no error")
except Exception as e:
    print(kw, "->", str(e)[:300])
