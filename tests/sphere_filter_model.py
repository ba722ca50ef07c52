"""TEST INFRASTRUCTURE: a numpy model of the decision logic of sphere_nn_filter_kernel
(vif_b200/csrc/xm_sphere.cu) -- same grid, same gaps, same culling mask, same completeness and
ambiguity tests, formula for formula -- so that the RIGOUR of those bounds can be checked on the CPU
against a brute-force search over many random geometries (tests/test_sphere_filter_model.py):
whenever the model accepts a row, its winner must be the true nearest neighbour. It is not a CPU
implementation of the product: it only says accept / flag and which slot won."""
import math

import numpy as np

D2R = 3.14159265358979323/180.0
K_INF_HI = 0x7ff00000
# cell order of the kernel: own, N, S, E, W, NE, NW, SE, SW
NB = [(0, 0), (0, 1), (0, -1), (1, 0), (-1, 0), (1, 1), (-1, 1), (1, -1), (-1, -1)]


def sphere_grid(bb, n_cand, nth=1, per_cell=8.0):
    """sphere_grid() of xm_sphere.cu."""
    pi = 3.141592653589793
    ra0, ra1, de0, de1 = bb
    de0, de1 = max(de0, -90.0), min(de1, 90.0)
    wrap = (ra1 - ra0) > 180.0
    span = 360.0 if wrap else max(ra1 - ra0, 1e-9)
    omega = (span*pi/180.0)*max(math.sin(de1*pi/180.0) - math.sin(de0*pi/180.0), 1e-12)
    h = math.sqrt(omega*per_cell*nth/max(n_cand, 1))*180.0/pi
    h = min(max(h, 1.0/3600.0), 10.0)
    nrow = math.ceil((de1 - de0)/h) + 1.0
    ncol = math.floor(360.0/h) if wrap else math.ceil(span/h) + 1.0
    ncol = max(ncol, 1.0)
    g = dict(nra=int(ncol), ndec=int(nrow), dra=(360.0/ncol if wrap else h), ddec=h, rra0=ra0,
             rdec0=de0 - 0.5*(nrow*h - (de1 - de0)), wrap=wrap)
    unit = pi/180.0
    g.update(ox=g["rra0"]*unit, oy=g["rdec0"]*unit, wx=g["dra"]*unit, wy=g["ddec"]*unit,
             fzx=1e-12*max(abs(ra0), abs(ra1), 360.0)*unit, fzy=1e-12*90.0*unit)
    return g


def cell_keys(ra, dec, g):
    """cell_key() of xm_bucket.cu (degrees, clamped)."""
    ix = np.clip(np.floor((ra - g["rra0"])/g["dra"]), 0, g["nra"] - 1).astype(np.int64)
    iy = np.clip(np.floor((dec - g["rdec0"])/g["ddec"]), 0, g["ndec"] - 1).astype(np.int64)
    return ix, iy


def cmin_rows(g):
    """sphere_rows_kernel."""
    r = np.arange(g["ndec"], dtype=np.float64)
    lo = g["oy"] + r*g["wy"] - g["fzy"]
    hi = lo + g["wy"] + 2.0*g["fzy"]
    v = np.minimum(np.cos(lo), np.cos(hi))*(1.0 - 1e-12) - 1e-15
    half_pi = 1.5707963267948966
    v = np.where((lo > -half_pi) & (hi < half_pi), np.maximum(v, 0.0), 0.0)
    return v


def _hi(q):
    return (np.asarray(q, dtype=np.float64).view(np.int64) >> 32).astype(np.int64)


def _hi_floor(h):
    return np.array([int(h) << 32], dtype=np.int64).view(np.float64)[0]


def _hi_ceil(h):
    return _hi_floor(K_INF_HI if h >= K_INF_HI else h + 1)


def filter_applicable(g):
    wmax = max(g["wx"], g["wy"])
    return wmax <= 0.02 and g["nra"] >= 8 and g["ndec"] >= 3


def run_filter(ra1, dec1, ra2, dec2, self_match=False):
    """-> (accepted mask over catalog-1 rows, winner row in catalog 2 for accepted rows, grid)."""
    bb = (min(ra1.min(), ra2.min()), max(ra1.max(), ra2.max()), min(dec1.min(), dec2.min()), max(dec1.max(), dec2.max()))
    g = sphere_grid(bb, ra2.size)
    assert filter_applicable(g), g
    wx, wy = g["wx"], g["wy"]
    wmax = max(wx, wy)
    eps_lo = (2.0*wmax)*(2.0*wmax)/12.0 + 1e-14
    shrink = 1.0 - eps_lo - 1e-9
    cmin = cmin_rows(g)
    ncol, nrow = g["nra"], g["ndec"]
    # candidates grouped by cell
    cx, cy = cell_keys(ra2, dec2, g)
    ck = cx*nrow + cy
    order = np.argsort(ck, kind="stable")
    cks = ck[order]
    cstart = np.searchsorted(cks, np.arange(ncol*nrow + 1))
    X2, Y2 = (ra2*D2R)[order], (dec2*D2R)[order]
    C2 = np.cos(Y2)
    sx, sy = cell_keys(ra1, dec1, g)
    X1, Y1 = ra1*D2R, dec1*D2R
    C1 = np.cos(Y1)
    accepted = np.zeros(ra1.size, dtype=bool)
    winner = np.full(ra1.size, -1, dtype=np.int64)
    for i in range(ra1.size):
        x0, y0, x1, y1, c1 = int(sx[i]), int(sy[i]), X1[i], Y1[i], C1[i]
        if not (c1 >= 0.0) or y0 < 1 or y0 + 1 >= nrow or x0 < 1 or x0 + 1 >= ncol:
            continue
        cx_lo, cy_lo = g["ox"] + x0*wx, g["oy"] + y0*wy
        ge = max(cx_lo + wx - x1 - g["fzx"], 0.0); gw = max(x1 - cx_lo - g["fzx"], 0.0)
        gn = max(cy_lo + wy - y1 - g["fzy"], 0.0); gs = max(y1 - cy_lo - g["fzy"], 0.0)
        cm0, cmn, cms = cmin[y0]*c1, cmin[y0 + 1]*c1, cmin[y0 - 1]*c1
        h1, h2, slot = K_INF_HI, K_INF_HI, -1

        def scan(k, h1, h2, slot):
            qs, qe = cstart[k], cstart[k + 1]
            if qe == qs:
                return h1, h2, slot
            dx, dy = X2[qs:qe] - x1, Y2[qs:qe] - y1
            h = _hi(dy*dy + dx*dx*C2[qs:qe]*c1)
            if self_match:
                h = np.where(order[qs:qe] == i, K_INF_HI, h)
            for t in range(h.size):                     # the kernel's running update, in slot order
                better = h[t] < h1
                h2 = min(h2, max(h1, int(h[t])))
                h1 = min(h1, int(h[t]))
                if better:
                    slot = qs + t
            return h1, h2, slot

        key = x0*nrow + y0
        h1, h2, slot = scan(key, h1, h2, slot)                             # A. own cell
        thr = _hi_ceil(h1)*(1.0 + 1e-9)                                    # B. culling mask
        todo = []
        for e in range(1, 9):
            bx, by = NB[e]
            gx = ge if bx > 0 else (gw if bx < 0 else 0.0)
            gy = gn if by > 0 else (gs if by < 0 else 0.0)
            cm = cmn if by > 0 else (cms if by < 0 else cm0)
            if (gx*gx*cm + gy*gy)*shrink < thr:
                todo.append(e)
        for e in todo:                                                     # C. surviving neighbours
            bx, by = NB[e]
            h1, h2, slot = scan(key + bx*nrow + by, h1, h2, slot)
        oy, ox = wy - 2.0*g["fzy"], wx - 2.0*g["fzx"]
        hy = 0.5*(min(gn, gs) + oy)
        hx = 0.5*min(min(ge, gw) + ox, 3.141592653589793)
        cmin3 = min(cm0, cmn, cms)
        out_p = min(hy*hy*(1.0 - hy*hy/3.0), hx*hx*(1.0 - hx*hx/3.0)*cmin3)*(1.0 - 1e-9)
        q1 = _hi_ceil(h1)
        flag = not (0.25*q1*(1.0 + 1e-9) <= out_p)
        if h2 < K_INF_HI and not (q1*(1.0 + 1e-12) < _hi_floor(h2)*shrink):
            flag = True
        if slot < 0 or h1 < 0:
            flag = True
        if not flag and not (abs(X2[slot] - x1) < 0.25 and abs(Y2[slot] - y1) < 0.25):
            flag = True
        if not flag:
            accepted[i] = True
            winner[i] = order[slot]
    return accepted, winner, g


def brute_nearest(ra1, dec1, ra2, dec2, self_match=False, chunk=512):
    """Reference proxy (qxmatch.hpp:192-204) over all pairs: nearest row, its proxy, the runner-up's proxy."""
    X1, Y1, X2, Y2 = ra1*D2R, dec1*D2R, ra2*D2R, dec2*D2R
    C1, C2 = np.cos(Y1), np.cos(Y2)
    best = np.empty(ra1.size, dtype=np.int64)
    p1 = np.empty(ra1.size)
    p2 = np.empty(ra1.size)
    for s in range(0, ra1.size, chunk):
        e = min(s + chunk, ra1.size)
        sra = np.sin(0.5*(X2[None, :] - X1[s:e, None]))
        sde = np.sin(0.5*(Y2[None, :] - Y1[s:e, None]))
        p = sde*sde + sra*sra*C2[None, :]*C1[s:e, None]
        if self_match:
            p[np.arange(e - s), np.arange(s, e)] = np.inf
        part = np.argpartition(p, 1, axis=1)[:, :2]
        rows = np.arange(e - s)
        a, b = p[rows, part[:, 0]], p[rows, part[:, 1]]
        first = np.where(a <= b, part[:, 0], part[:, 1])
        best[s:e] = first
        p1[s:e] = np.minimum(a, b)
        p2[s:e] = np.maximum(a, b)
    return best, p1, p2
