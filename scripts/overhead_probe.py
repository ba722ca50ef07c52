import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes as C
from vif_b200 import qxmatch, catalogs as G, _lib
from vif_b200.qxmatch import QxmatchParams, last_stats
n = 10_000_000
a1, b1 = G.c2(1, n, xp="torch", device="cuda"); a2, b2 = G.c2(2, n, xp="torch", device="cuda")
for _ in range(3): r = qxmatch(a1, b1, a2, b2)
torch.cuda.synchronize()
for it in range(4):
    t0 = time.perf_counter(); r = qxmatch(a1, b1, a2, b2); t1 = time.perf_counter()
    print(f"qxmatch() wall {1e3*(t1-t0):.3f} ms; ms_host {r.stats['ms_host']:.3f}; ms_total {r.stats['ms_total']:.3f}")
lib = _lib.load()
ident = torch.empty((1, n), dtype=torch.int64, device="cuda"); d = torch.empty((1, n), dtype=torch.float64, device="cuda")
rid = torch.empty(n, dtype=torch.int64, device="cuda"); rd = torch.empty(n, dtype=torch.float64, device="cuda")
p = QxmatchParams(device=0).to_c(); err = C.create_string_buffer(512)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = lib.xm_qxmatch_dev(a1.data_ptr(), b1.data_ptr(), n, a2.data_ptr(), b2.data_ptr(), n, C.byref(p), ident.data_ptr(), d.data_ptr(), rid.data_ptr(), rd.data_ptr(), None, err, 512)
    t1 = time.perf_counter()
    print(f"raw xm_qxmatch_dev wall {1e3*(t1-t0):.3f} ms rc {rc}; ms_host {last_stats(0)['ms_host']:.3f}")
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x = [torch.empty((1, n), dtype=torch.int64, device="cuda") for _ in range(4)]
    t1 = time.perf_counter(); print(f"4x torch.empty {1e3*(t1-t0):.3f} ms")
