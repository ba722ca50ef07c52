"""CPU-only randomized sweeps of the two filter-logic models (tests/ring_filter_model.py against the
oracle, tests/sphere_filter_model.py against brute force): every row a model accepts must carry the
reference's id / the true nearest neighbour. usage: python scripts/model_sweep.py [ring|sphere] [configs] [seed]
Round-1 runs: ring 150 configs seed 12345 -> 123 applicable, 283 374 rows, 283 364 accepted (55 472 stopped
after ring 0), 0 wrong; sphere 120 configs seed 777 -> 101 applicable, 153 989 rows, 104 522 accepted, 0 wrong;
ring 400 configs seed 2024 -> 326 applicable, 725 168 rows, 725 140 accepted, 0 wrong; sphere 300 configs seed 2025
-> 247 applicable, 365 015 rows, 220 127 accepted, 0 wrong."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def ring(nconf, seed):
    import oracle as O
    from ring_filter_model import run_filter, grid_view, fast_path_applicable
    rng = np.random.default_rng(seed)
    tot = acc_tot = wrong_tot = skipped = ring0_tot = 0
    for t in range(nconf):
        dec_c, ra_c = rng.uniform(-70, 70), rng.uniform(0, 350)
        hd, hr = 10**rng.uniform(-1.5, 0.8), 10**rng.uniform(-1.5, 0.8)
        d0, d1 = max(dec_c - hd, -85), min(dec_c + hd, 85)
        r0, r1 = ra_c, ra_c + 2*hr
        n1, n2 = int(10**rng.uniform(2.5, 3.9)), int(10**rng.uniform(2.0, 4.3))
        self_match = t % 7 == 0
        ra1, dec1 = rng.uniform(r0, r1, n1), rng.uniform(d0, d1, n1)
        if self_match:
            ra2, dec2 = ra1, dec1
        else:
            ra2, dec2 = rng.uniform(r0, r1, n2), rng.uniform(d0, d1, n2)
            if t % 5 == 0:      # displaced copies of a subset
                m = min(n1, n2)//2
                ra2[:m] = ra1[:m] + rng.normal(0, 0.3/3600, m)
                dec2[:m] = dec1[:m] + rng.normal(0, 0.3/3600, m)
        bb1 = (ra1.min(), ra1.max(), dec1.min(), dec1.max())
        bb2 = (ra2.min(), ra2.max(), dec2.min(), dec2.max())
        G = O.oracle_grid(bb1, bb2, ra2.size)
        if G["nra"]*G["ndec"] > 3e6 or not fast_path_applicable(grid_view(G)):
            skipped += 1
            continue
        acc, win, ring0 = run_filter(ra1, dec1, ra2, dec2, G, self_match)
        ref = O.oracle_qxmatch(ra1, dec1, ra2, dec2, self=self_match, thread=os.cpu_count() or 8,
                               no_mirror=True)["id"][0].astype(np.int64)
        wrong = np.flatnonzero(acc & (win != ref))
        tot += n1; acc_tot += int(acc.sum()); wrong_tot += wrong.size; ring0_tot += int(ring0.sum())
        if wrong.size:
            print("WRONG", t, wrong[:5], win[wrong[:5]], ref[wrong[:5]], G, flush=True)
    print("ring: configs", nconf - skipped, "rows", tot, "accepted", acc_tot, "ring-0 stops", ring0_tot, "wrong", wrong_tot)
    return wrong_tot


def sphere(nconf, seed):
    from sphere_filter_model import run_filter, brute_nearest, sphere_grid, filter_applicable
    rng = np.random.default_rng(seed)
    tot = acc_tot = wrong = skipped = 0
    for t in range(nconf):
        kind = t % 4
        n1, n2 = int(10**rng.uniform(2.5, 3.6)), int(10**rng.uniform(2.8, 4.2))
        if kind == 0:       # polar cap
            rad, south = 10**rng.uniform(-0.5, 1.0), rng.random() < 0.5

            def gen(n):
                ra, dec = rng.uniform(0, 360, n), 90 - rad*np.sqrt(rng.uniform(0, 1, n))
                return ra, (-dec if south else dec)
        elif kind == 1:     # band around the whole circle
            d, hw = rng.uniform(-80, 80), 10**rng.uniform(-0.7, 0.5)

            def gen(n):
                return rng.uniform(0, 360, n), np.clip(rng.uniform(d - hw, d + hw, n), -89.99, 89.99)
        elif kind == 2:     # box anywhere
            d, hw, r, rw = rng.uniform(-88, 88), 10**rng.uniform(-1, 0.7), rng.uniform(0, 300), 10**rng.uniform(-1, 1.5)

            def gen(n):
                return rng.uniform(r, r + rw, n), np.clip(rng.uniform(d - hw, d + hw, n), -89.999, 89.999)
        else:               # clumps
            d, r = rng.uniform(-60, 60), rng.uniform(0, 300)
            cra, cde = rng.uniform(r, r + 5, 30), rng.uniform(d - 2, d + 2, 30)

            def gen(n):
                w = rng.integers(0, 30, n)
                return cra[w] + rng.normal(0, 0.03, n), np.clip(cde[w] + rng.normal(0, 0.03, n), -89.9, 89.9)
        c1 = gen(n1)
        selfm = t % 9 == 0
        c2 = c1 if selfm else gen(n2)
        bb = (min(c1[0].min(), c2[0].min()), max(c1[0].max(), c2[0].max()),
              min(c1[1].min(), c2[1].min()), max(c1[1].max(), c2[1].max()))
        g = sphere_grid(bb, c2[0].size)
        if not filter_applicable(g) or g["nra"]*g["ndec"] > 5e6:
            skipped += 1
            continue
        acc, win, _ = run_filter(c1[0], c1[1], c2[0], c2[1], selfm)
        best, p1, p2 = brute_nearest(c1[0], c1[1], c2[0], c2[1], selfm)
        w = np.flatnonzero(acc & (win != best))
        tot += c1[0].size; acc_tot += int(acc.sum()); wrong += w.size
        if w.size:
            print("WRONG", t, kind, w[:5], win[w[:5]], best[w[:5]], p1[w[:5]], p2[w[:5]], g, flush=True)
    print("sphere: configs", nconf - skipped, "rows", tot, "accepted", acc_tot, "wrong", wrong)
    return wrong


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "ring"
    nconf = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    sys.exit(1 if (ring if which == "ring" else sphere)(nconf, seed) else 0)
