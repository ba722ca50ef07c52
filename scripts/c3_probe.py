"""C3-shaped probe: self match nth=5 on the clustered generator; parity at 1e6 vs the oracle, timing at 1e7/1e8."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from vif_b200 import qxmatch, catalogs as G
lv = [int(a) for a in sys.argv[1:]] or [6, 7]
for L in lv:
    ra, dec = G.c3_clustered(3, L, xp="torch", device="cuda")
    torch.cuda.synchronize()
    for it in range(2):
        t = time.time(); r = qxmatch(ra, dec, nth=5); torch.cuda.synchronize(); w = time.time() - t
    st = r.stats
    n = ra.numel()
    dup = int((r.id[0] == r.id[1]).sum())
    print(f"C3 levels={L} n={n}: wall {w*1e3:.1f} ms total {st['ms_total']:.2f} (index {st['ms_index']:.2f} fwd {st['ms_forward']:.2f}) -> {n/(st['ms_total']*1e-3):.3e} src/s; "
          f"evals/src {st['evals_fwd']/n:.0f}; dup-id rows {dup} ({100*dup/n:.1f}%); grid {st['grid']['nra']}x{st['grid']['ndec']} cs {st['grid']['cell_size']:.2f}; ambiguous {st['rows_ambiguous']}", flush=True)
    if L <= 6:
        a, b = ra.cpu().numpy(), dec.cpu().numpy()
        t = time.time(); exp = O.oracle_qxmatch(a, b, a, b, nth=5, self=True, thread=os.cpu_count()); to = time.time() - t
        ok = np.array_equal(exp["id"], r.id.cpu().numpy().view(np.uint64))
        dd = r.d.cpu().numpy(); f = np.isfinite(exp["d"])
        rel = np.max(np.abs(dd[f]-exp["d"][f])/np.maximum(exp["d"][f], 1e-300))
        print(f"   parity vs oracle ({to:.1f}s on {os.cpu_count()} threads): ids {'OK' if ok else 'FAIL'} maxrel d {rel:.2e}", flush=True)
