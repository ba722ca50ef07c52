"""Brute-force probe: parity at 5e4 x 5e4 vs the oracle, timing at C4 size (1e6 x 1e6 full sky)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from vif_b200 import qxmatch, catalogs as G

def cmp(tag, got, exp):
    ok = all(np.array_equal(getattr(got, k), exp[k]) for k in ("id", "rid"))
    rel = 0.0
    for k in ("d", "rd"):
        a, b = getattr(got, k), exp[k]
        f = np.isfinite(b)
        if f.any(): rel = max(rel, np.max(np.abs(a[f]-b[f])/np.maximum(b[f], 1e-300)))
        ok &= np.array_equal(np.isnan(a), np.isnan(b))
    st = got.stats
    print(f"[{'OK' if ok and rel < 1e-9 else 'FAIL'}] {tag}: maxrel {rel:.2e} total {st['ms_total']:.2f} ms fwd {st['ms_forward']:.2f} rev {st['ms_reverse']:.2f} refined {st['rows_refined_fwd']}/{st['rows_refined_rev']}")

n = 50000
a1, b1 = G.fullsky(1, n); a2, b2 = G.fullsky(2, n)
for kw in (dict(), dict(nth=3), dict(nth=2, self=True)):
    same = kw.get("self")
    exp = O.oracle_qxmatch(a1, b1, a1 if same else a2, b1 if same else b2, brute_force=True, thread=os.cpu_count(), **kw)
    got = qxmatch(a1, b1, None if same else a2, None if same else b2, brute_force=True, **{k: v for k, v in kw.items() if k != "self"})
    cmp(f"5e4 {kw}", got, exp)
# duplicates -> many flagged rows
d2a, d2b = a2.copy(), b2.copy(); d2a[n//2:] = d2a[:n//2]; d2b[n//2:] = d2b[:n//2]
exp = O.oracle_qxmatch(a1[:20000], b1[:20000], d2a, d2b, brute_force=True, nth=2, thread=os.cpu_count())
cmp("dups nth2", qxmatch(a1[:20000], b1[:20000], d2a, d2b, brute_force=True, nth=2), exp)
# C4 timing
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
r1, c1 = G.fullsky(1, N, xp="torch", device="cuda"); r2, c2 = G.fullsky(2, N, xp="torch", device="cuda")
for it in range(2):
    t = time.time(); res = qxmatch(r1, c1, r2, c2, brute_force=True); torch.cuda.synchronize(); w = time.time() - t
    st = res.stats
    pairs = N*N
    print(f"C4 {N}x{N}: wall {w:.3f}s total {st['ms_total']:.1f} ms fwd {st['ms_forward']:.1f} rev {st['ms_reverse']:.1f} index {st['ms_index']:.2f} "
          f"-> {N/(st['ms_total']*1e-3):.3e} src/s, fwd {pairs/(st['ms_forward']*1e-3):.3e} pairs/s = {11*pairs/(st['ms_forward']*1e-3)/1e12:.1f} TFLOP/s(11-flop conv) refined {st['rows_refined_fwd']}/{st['rows_refined_rev']}")
sel = np.arange(0, N, N//48)[:48]
h1, k1, h2, k2 = (t.cpu().numpy() for t in (r1, c1, r2, c2))
exp = O.oracle_qxmatch(h1[sel], k1[sel], h2, k2, brute_force=True, no_mirror=True, thread=os.cpu_count())
print("sampled fwd ids equal:", np.array_equal(exp["id"][0], res.id[0][torch.from_numpy(sel).cuda()].cpu().numpy().view(np.uint64)))
exp = O.oracle_qxmatch(h2[sel], k2[sel], h1, k1, brute_force=True, no_mirror=True, thread=os.cpu_count())
print("sampled rev ids equal:", np.array_equal(exp["id"][0], res.rid[torch.from_numpy(sel).cuda()].cpu().numpy().view(np.uint64)))
