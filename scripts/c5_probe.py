"""C5-shaped probe: full-sky n1 x n2, nth=1, binned call (-> exact spherical grid search); sampled rows
checked against the oracle's brute force."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle as O
from vif_b200 import qxmatch, catalogs as G
n1 = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
n2 = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
def gen(seed, n):   # chunked to bound temporaries
    ra = torch.empty(n, dtype=torch.float64, device="cuda"); dec = torch.empty(n, dtype=torch.float64, device="cuda")
    step = 50_000_000
    for s in range(0, n, step):
        m = min(step, n - s)
        a, b = G.fullsky(seed, m, xp="torch", device="cuda", start=s)
        ra[s:s+m] = a; dec[s:s+m] = b
    return ra, dec
ra1, dec1 = gen(1, n1); ra2, dec2 = gen(2, n2)
torch.cuda.synchronize()
print("generated", n1, n2, "free GB", torch.cuda.mem_get_info()[0]/1e9, flush=True)
for it in range(int(os.environ.get("ITERS", "2"))):
    t = time.time(); r = qxmatch(ra1, dec1, ra2, dec2); torch.cuda.synchronize(); w = time.time() - t
    st = r.stats
    print(f"C5 {n1:.0e}x{n2:.0e}: wall {w:.3f}s total {st['ms_total']:.1f} ms (index {st['ms_index']:.1f} fwd {st['ms_forward']:.1f} rev {st['ms_reverse']:.1f}) "
          f"-> {n1/(st['ms_total']*1e-3):.3e} src/s; evals/src fwd {st['evals_fwd']/n1:.1f} rev {st['evals_rev']/n2:.1f}; grid {st['grid']['nra']}x{st['grid']['ndec']} cell {st["grid"]["cell_size"]:.1f}\" refined {st["rows_refined_fwd"]}/{st["rows_refined_rev"]}", flush=True)
sel = torch.arange(0, n1, n1//48, device="cuda")[:48]
h2, k2 = ra2.cpu().numpy(), dec2.cpu().numpy()
exp = O.oracle_qxmatch(ra1[sel].cpu().numpy(), dec1[sel].cpu().numpy(), h2, k2, brute_force=True, no_mirror=True, thread=os.cpu_count())
print("sampled forward ids equal:", np.array_equal(exp["id"][0], r.id[0][sel].cpu().numpy().view(np.uint64)),
      "max rel d", float(np.max(np.abs(exp["d"][0]-r.d[0][sel].cpu().numpy())/exp["d"][0])), flush=True)
if n1 <= 200_000_000:
    sel2 = torch.arange(0, n2, n2//16, device="cuda")[:16]
    h1, k1 = ra1.cpu().numpy(), dec1.cpu().numpy()
    exp = O.oracle_qxmatch(ra2[sel2].cpu().numpy(), dec2[sel2].cpu().numpy(), h1, k1, brute_force=True, no_mirror=True, thread=os.cpu_count())
    print("sampled reverse ids equal:", np.array_equal(exp["id"][0], r.rid[sel2].cpu().numpy().view(np.uint64)), flush=True)
