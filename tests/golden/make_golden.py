"""Mints tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, i.e. vif::astro::qxmatch
compiled in place from /root/reference). Run in the authoring container only:

    python tests/golden/make_golden.py

The reference ships no test for this path, so these fixtures ARE the pin:
the oracle port, the host geometry and the CUDA path are all checked against them. Each file
holds the inputs, the keyword set and the reference outputs (id, d, rid, rd); geometry.npz holds
field_area_hull values and ring tables dumped from the reference's depth_cache.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402
from vif_b200 import catalogs as G  # noqa: E402


def save(name, ra1, dec1, ra2, dec2, **kw):
    r = O.ref_qxmatch(ra1, dec1, ra2, dec2, **kw)
    keys = {"kw_" + k: np.array(v) for k, v in kw.items()}
    np.savez_compressed(os.path.join(HERE, name + ".npz"), ra1=ra1, dec1=dec1, ra2=ra2, dec2=dec2,
                        id=r["id"], d=r["d"], rid=r["rid"], rd=r["rd"], **keys)
    print(name, {k: v.shape for k, v in r.items()})


def main():
    a1, b1 = G.c1(1, 1500)
    a2, b2 = G.c1(2, 1200)
    save("box_nth1", a1, b1, a2, b2)
    save("box_nth2", a1, b1, a2, b2, nth=2)
    save("box_nth3_thread3", a1, b1, a2, b2, nth=3, thread=3)
    save("box_no_mirror", a1, b1, a2, b2, no_mirror=True)
    save("box_self_nth5", a1, b1, a1, b1, nth=5, self=True)
    save("box_nth12", a1[:400], b1[:400], a2[:300], b2[:300], nth=12)
    save("box_linear", a1[:500], b1[:500], a2[:400], b2[:400], linear=True)
    save("box_linear_nth2_brute", a1[:500], b1[:500], a2[:400], b2[:400], linear=True, nth=2, brute_force=True)
    # n2 < nth: bucket path duplicates ids, brute force does not
    s1, t1 = G.box(3, 200, 0, 5, 0, 5)
    s2, t2 = G.box(4, 3, 0, 5, 0, 5)
    save("few_n2_3_nth5", s1, t1, s2, t2, nth=5)
    save("few_n2_3_nth5_brute", s1, t1, s2, t2, nth=5, brute_force=True)
    save("few_n2_1", *G.box(5, 40, 0, 0.03, 0, 0.03), *G.box(6, 1, 0, 0.03, 0, 0.03))
    save("few_n1_1", *G.box(6, 1, 0, 0.03, 0, 0.03), *G.box(5, 40, 0, 0.03, 0, 0.03), nth=2)
    # exact ties: duplicated points in both catalogs
    d1a, d1b = G.c1(7, 600)
    d2a, d2b = G.c1(8, 500)
    d2a[250:] = d2a[:250]; d2b[250:] = d2b[:250]
    d1a[300:] = d1a[:300]; d1b[300:] = d1b[:300]
    save("ties_nth1", d1a, d1b, d2a, d2b)
    save("ties_nth3", d1a, d1b, d2a, d2b, nth=3)
    save("ties_brute_nth2", d1a, d1b, d2a, d2b, nth=2, brute_force=True)
    save("ties_self_nth2", d2a, d2b, d2a, d2b, nth=2, self=True)
    # cat-1 identical to cat-2 without `self`: zero distances
    save("same_points_noself", d2a[:250], d2b[:250], d2a[:250], d2b[:250])
    # lattices: many symmetric configurations
    gx, gy = np.meshgrid(np.arange(24) / 64.0, np.arange(24) / 64.0, indexing="ij")
    lx, ly = gx.ravel() + 10.0, gy.ravel() + 20.0
    save("lattice_nth4", lx + 1.0 / 128, ly + 1.0 / 128, lx, ly, nth=4)
    save("lattice_self_nth3", lx, ly, lx, ly, nth=3, self=True)
    # other fields: high declination, RA near 360, tall field (reference is inexact there), pole
    save("dec60_70", *G.box(9, 1500, 0, 10, 60, 10), *G.box(10, 1500, 0, 10, 60, 10))
    save("ra350_360_nth2", *G.box(11, 1200, 350, 10, -5, 10), *G.box(12, 900, 350, 10, -5, 10), nth=2)
    save("tall_dec40_80", *G.box(13, 2000, 0, 10, 40, 40), *G.box(14, 2000, 0, 10, 40, 40))
    save("pole_dec80_899", *G.box(15, 2000, 0, 10, 80, 9.9), *G.box(16, 2000, 0, 10, 80, 9.9), nth=2)
    save("south_dec_m70_m60", *G.box(17, 800, 100, 3, -70, 10), *G.box(18, 1000, 100, 3, -70, 10))
    # different extents of the two catalogs (grid from cat-2's box, bounds from both)
    save("offset_boxes", *G.box(19, 1000, 0, 2, 0, 2), *G.box(20, 1000, 0.5, 1, 0.5, 1), nth=2)
    # clustered
    cr, cd = G.c3_clustered(3, 3)
    save("clustered_self_nth5", cr, cd, cr, cd, nth=5, self=True)
    # brute force, full sky, reverse match
    f1 = G.fullsky(21, 800)
    f2 = G.fullsky(22, 700)
    save("fullsky_brute", *f1, *f2, brute_force=True)
    save("fullsky_brute_nth4", *f1, *f2, brute_force=True, nth=4)
    save("fullsky_brute_self_nth2", *f1, *f1, brute_force=True, nth=2, self=True)
    save("fullsky_brute_no_mirror", *f1, *f2, brute_force=True, no_mirror=True, nth=9)
    # empty catalogs (qxmatch.hpp:165-167)
    e = np.zeros(0)
    save("empty_cat2", a1[:10], b1[:10], e, e, nth=2)
    save("empty_cat1", e, e, a2[:10], b2[:10])

    # geometry: hull areas + ring tables straight from the reference
    boxes = np.array([[0, 1, 0, 1], [0, 10, -5, 5], [0, 10, 60, 70], [350, 360, -5, 5], [0, 10, 40, 80],
                      [0, 10, 80, 89.9], [0, 360, -90, 90], [3, 3, 1, 2], [3, 4, 2, 2], [5, 5, 5, 5],
                      [100, 103, -70, -60], [0.5, 1.5, 0.5, 1.5], [-3, 2, -1, 4]], dtype=np.float64)
    areas = np.array([O.ref_field_area_bbox(*b) for b in boxes])
    grids = np.array([[91, 91], [5, 40], [3, 3], [2, 7], [200, 17], [12, 1000]], dtype=np.int64)
    rings = {}
    for nx, ny in grids:
        for k in range(0, min(max(nx, ny), 14)):
            bx, by, md = O.ref_ring(nx, ny, 40.0, k)
            rings[f"ring_{nx}_{ny}_{k}_bx"] = bx.astype(np.int32)
            rings[f"ring_{nx}_{ny}_{k}_by"] = by.astype(np.int32)
            rings[f"ring_{nx}_{ny}_{k}_md"] = np.array(md)
    angs = np.array([[0.1, 0.2, 0.3, 0.4], [10, -5, 10.01, -5.01], [359.9, 89, 0.1, 89.5], [0, 0, 180, 0],
                     [12.5, 45, 12.5, 45]], dtype=np.float64)
    angd = np.array([O.ref_angdist(*a) for a in angs])
    np.savez_compressed(os.path.join(HERE, "geometry.npz"), boxes=boxes, areas=areas, grids=grids,
                        angs=angs, angd=angd, **rings)
    print("geometry", areas)


if __name__ == "__main__":
    main()
