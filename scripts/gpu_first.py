"""First-contact GPU script: runs the CUDA path against the CPU oracle on a batch of cases and prints
mismatch statistics + stage timings. Diagnostic only (tests/ holds the real parity tests)."""
import sys, time, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle as O
from vif_b200 import qxmatch, catalogs as G, _lib
import ctypes as C


def compare(tag, got, exp, rtol=1e-9):
    out = []
    ok = True
    for k in ("id", "rid"):
        a, b = getattr(got, k), exp[k]
        if a.shape != b.shape:
            out.append(f"{k}: shape {a.shape} vs {b.shape}"); ok = False; continue
        nd = int(np.sum(a != b))
        if nd: ok = False
        out.append(f"{k}: {nd} diff / {a.size}")
    for k in ("d", "rd"):
        a, b = getattr(got, k), exp[k]
        if a.shape != b.shape:
            out.append(f"{k}: shape"); ok = False; continue
        nanm = int(np.sum(np.isnan(a) != np.isnan(b)))
        infm = int(np.sum(np.isinf(a) != np.isinf(b)))
        fin = np.isfinite(a) & np.isfinite(b)
        rel = np.max(np.abs(a[fin] - b[fin]) / np.maximum(np.abs(b[fin]), 1e-300)) if fin.any() else 0.0
        nbit = int(np.sum(a[fin] != b[fin]))
        if nanm or infm or rel > rtol: ok = False
        out.append(f"{k}: maxrel {rel:.2e} notbitequal {nbit} nan-mismatch {nanm} inf-mismatch {infm}")
    st = got.stats
    print(f"[{'OK' if ok else 'FAIL'}] {tag}: " + "; ".join(out))
    print(f"       refined {st['rows_refined_fwd']}/{st['rows_refined_rev']} host {st['ms_host']:.3f} total {st['ms_total']:.3f} ms (bbox {st['ms_bbox']:.3f} keys {st['ms_keys']:.3f} search {st['ms_search']:.3f} "
          f"index {st['ms_index']:.3f} fwd {st['ms_forward']:.3f} rev {st['ms_reverse']:.3f}) launches {st['launches']} "
          f"evals {st['evals_fwd']}/{st['evals_rev']} ambiguous {st['rows_ambiguous']} sin_model {st['sin_model']} "
          f"grid {st['grid']['nra']}x{st['grid']['ndec']} cs {st['grid']['cell_size']:.4f}")
    return ok


def run(tag, ra1, dec1, ra2, dec2, othreads=8, **kw):
    try:
        t = time.time(); got = qxmatch(ra1, dec1, ra2, dec2, **kw); tg = time.time() - t
        t = time.time(); exp = O.oracle_qxmatch(ra1, dec1, ra2, dec2, thread=othreads, **{k: v for k, v in kw.items()}); to = time.time() - t
        ok = compare(f"{tag} {kw} gpu_wall {tg:.3f}s oracle {to:.2f}s", got, exp)
        return ok
    except Exception:
        traceback.print_exc()
        return False


def main():
    lib = _lib.load()
    print(lib.xm_version())
    # sort
    import torch
    for n, bits in [(1000, 7), (100000, 20), (3_000_000, 21), (4097, 32), (1, 5)]:
        k = torch.randint(0, 2**bits, (n,), dtype=torch.int64, device="cuda").to(torch.int32)
        k = (k.to(torch.int64) & 0xffffffff)
        ku = k.clone()
        keys = ku.to(torch.int32) if bits < 32 else (ku - (ku >= 2**31) * 2**32).to(torch.int32)
        vals = torch.arange(n, dtype=torch.int32, device="cuda")
        err = C.create_string_buffer(512)
        torch.cuda.synchronize()
        rc = lib.xm_sort_pairs_dev(keys.data_ptr(), vals.data_ptr(), n, bits, -1, None, err, 512)
        sk, order = torch.sort(k, stable=True)
        got_k = keys.to(torch.int64) & 0xffffffff
        print("sort", n, bits, "rc", rc, err.value, "keys ok", bool(torch.equal(got_k, sk)), "vals ok",
              bool(torch.equal(vals.to(torch.int64), order)))
    ok = True
    n = 20000
    SKIP = os.environ.get("SKIP_DONE") == "1"
    a1, b1 = G.c1(1, n); a2, b2 = G.c1(2, n)
    ok &= run("c1-20k", a1, b1, a2, b2)
    ok &= run("c1-20k", a1, b1, a2, b2, nth=2)
    ok &= run("c1-20k", a1, b1, a2, b2, nth=3, no_mirror=True)
    ok &= run("c1-20k-self5", a1, b1, a1, b1, nth=5, self=True)
    ok &= run("c1-20k-nth12", a1, b1, a2, b2, nth=12)
    ok &= run("c1-20k-linear", a1[:3000], b1[:3000], a2[:3000], b2[:3000], linear=True)
    ok &= run("c1-20k-linear-nth2", a1[:3000], b1[:3000], a2[:3000], b2[:3000], linear=True, nth=2)
    a1, b1 = G.c1(1, 100000); a2, b2 = G.c1(2, 100000)
    ok &= run("C1-full", a1, b1, a2, b2)
    a1, b1 = G.box(1, 50000, 0, 10, 60, 10); a2, b2 = G.box(2, 50000, 0, 10, 60, 10)
    ok &= run("dec60-70", a1, b1, a2, b2)
    a1, b1 = G.box(1, 50000, 350, 10, -5, 10); a2, b2 = G.box(2, 30000, 350, 10, -5, 10)
    ok &= run("ra350", a1, b1, a2, b2, nth=2)
    a1, b1 = G.box(1, 20000, 0, 10, 40, 40); a2, b2 = G.box(2, 20000, 0, 10, 40, 40)
    ok &= run("tall-dec40-80", a1, b1, a2, b2)
    a1, b1 = G.box(1, 2000, 0, 5, 0, 5); a2, b2 = G.box(2, 3, 0, 5, 0, 5)
    ok &= run("n2=3,nth=5", a1, b1, a2, b2, nth=5)
    a1, b1 = G.box(1, 50, 0, 0.05, 0, 0.05); a2, b2 = G.box(2, 1, 0, 0.05, 0, 0.05)
    ok &= run("n2=1", a1, b1, a2, b2)
    ra, dec = G.c3_clustered(3, 5)
    ok &= run("c3-1e5-self5", ra, dec, ra, dec, nth=5, self=True)
    # brute
    a1, b1 = G.fullsky(1, 3000); a2, b2 = G.fullsky(2, 2500)
    ok &= run("brute-fullsky", a1, b1, a2, b2, brute_force=True)
    ok &= run("brute-fullsky", a1, b1, a2, b2, brute_force=True, nth=4)
    ok &= run("brute-fullsky-self", a1, b1, a1, b1, brute_force=True, self=True, nth=2)
    ok &= run("brute-fullsky-nth10", a1, b1, a2, b2, brute_force=True, nth=10, no_mirror=True)
    ok &= run("brute-linear", a1, b1, a2, b2, brute_force=True, linear=True)
    # bigger: 1e6
    a1, b1 = G.c2(1, 1_000_000); a2, b2 = G.c2(2, 1_000_000)
    ok &= run("c2-1e6", a1, b1, a2, b2)
    print("ALL OK" if ok else "SOME FAILED")
    # timing at 1e7 (no oracle): device-resident
    a1, b1 = G.c2(1, 10_000_000, xp="torch", device="cuda"); a2, b2 = G.c2(2, 10_000_000, xp="torch", device="cuda")
    for it in range(3):
        r = qxmatch(a1, b1, a2, b2)
        st = r.stats
        print(f"C2 1e7 dev: host {st['ms_host']:.3f} refined {st['rows_refined_fwd']}/{st['rows_refined_rev']} total {st['ms_total']:.3f} ms (bbox {st['ms_bbox']:.3f} keys {st['ms_keys']:.3f} search {st['ms_search']:.3f} index {st['ms_index']:.3f} "
              f"fwd {st['ms_forward']:.3f} rev {st['ms_reverse']:.3f}) evals {st['evals_fwd']}/{st['evals_rev']} amb {st['rows_ambiguous']}")


if __name__ == "__main__":
    main()
