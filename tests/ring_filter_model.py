"""TEST INFRASTRUCTURE: a numpy model of the decision logic of ring_nn_filter_kernel
(vif_b200/csrc/xm_search_fast.cu + xm_filter_common.cuh + the GridView constants of xm_api.cu) --
the reference's grid, ring 0, the stop rule with its slack, the culling mask of the 5x5 block, the
stop rule after ring 1 and the ambiguity test, formula for formula. tests/test_ring_filter_model.py
runs it against the oracle (the reference's own binned search, quirks included): a row the filter
accepts must carry the id the reference returns. Not a CPU implementation of the product: it only
says accept / flag and which candidate won."""
import numpy as np

DPI = 3.14159265358979323
D2R = DPI/180.0
K_INF_HI = 0x7ff00000
BX = [0, 1, 0, -1, 1, -1, -1, 1, 0, 1, 2, 2, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1]
BY = [1, 0, -1, 0, 1, 1, -1, -1, 2, 2, 2, 1, 0, -1, -2, -2, -2, -2, -2, -1, 0, 1, 2, 2]


def _hi(q):
    return (np.asarray(q, dtype=np.float64).view(np.int64) >> 32).astype(np.int64)


def _hi_floor(h):
    return np.array([int(h) << 32], dtype=np.int64).view(np.float64)[0]


def _hi_ceil(h):
    return _hi_floor(K_INF_HI if h >= K_INF_HI else h + 1)


def grid_view(G):
    """GridView constants as run_device_once() sets them (xm_api.cu) + row_tables_kernel."""
    g = dict(G)
    g["ox"], g["oy"], g["wx"], g["wy"] = G["rra0"]*D2R, G["rdec0"]*D2R, G["dra"]*D2R, G["ddec"]*D2R
    mx = max(abs(G["rra0"]), abs(G["rra1"]), G["dra"])
    my = max(abs(G["rdec0"]), abs(G["rdec1"]), G["ddec"])
    g["fzx"], g["fzy"] = 1e-12*mx*D2R, 1e-12*my*D2R
    g["cs_rad"] = G["cell_size"]*(DPI/180.0/3600.0)
    H = 1.5*max(g["wx"], g["wy"])
    g["eps_lo"], g["eps_hi"] = H*H/3.0 + 1e-14, 1e-14
    g["stop_delta"] = g["cs_rad"]*(1e-9 + 2.0*H*H)
    cy = np.arange(int(G["ndec"]), dtype=np.float64)
    lo = g["oy"] + cy*g["wy"] - g["fzy"]
    hi = lo + g["wy"] + 2.0*g["fzy"]
    half_pi = 1.5707963267948966
    v = np.maximum(np.minimum(np.cos(lo), np.cos(hi))*(1.0 - 1e-12) - 1e-15, 0.0)
    g["cmin_row"] = np.where((lo >= -half_pi) & (hi <= half_pi), v, -1.0)
    g["ccen_row"] = np.cos(D2R*(G["rdec0"] + (cy + 0.5)*G["ddec"]))
    return g


def fast_path_applicable(g):
    half_pi = 1.5707963267948966
    if not (0 < g["wx"] <= 5e-3 and 0 < g["wy"] <= 5e-3):
        return False
    return g["oy"] >= -half_pi and g["oy"] + g["ndec"]*g["wy"] <= half_pi


def stop_decision(q1_lo, q1_hi, tr, g):
    lo, hi = tr - g["stop_delta"], tr + g["stop_delta"]
    if lo > 0.0 and q1_hi*(1.0 + g["eps_hi"]) <= lo*lo*(1.0 - lo*lo*(1.0/12.0) - 1e-12):
        return 0
    if hi <= 0.0:
        return 1 if q1_lo > 0.0 else 2
    if q1_lo*(1.0 - g["eps_lo"]) > hi*hi:
        return 1
    return 2


def run_filter(ra1, dec1, ra2, dec2, G, self_match=False):
    """G: the reference's grid (oracle_grid). -> (accepted mask, winner row of catalog 2, stopped-after-ring-0 mask)."""
    g = grid_view(G)
    assert fast_path_applicable(g), g
    nra, ndec = int(G["nra"]), int(G["ndec"])
    wx, wy = g["wx"], g["wy"]

    def keys(ra, dec):
        ix = np.clip(np.floor((ra - G["rra0"])/G["dra"]), 0, nra - 1).astype(np.int64)
        iy = np.clip(np.floor((dec - G["rdec0"])/G["ddec"]), 0, ndec - 1).astype(np.int64)
        return ix, iy

    cx, cy = keys(ra2, dec2)
    ck = cx*ndec + cy
    order = np.argsort(ck, kind="stable")
    cstart = np.searchsorted(ck[order], np.arange(nra*ndec + 1))
    X2, Y2 = (ra2*D2R)[order], (dec2*D2R)[order]
    C2 = np.cos(Y2)
    sx, sy = keys(ra1, dec1)
    X1, Y1 = ra1*D2R, dec1*D2R
    C1 = np.cos(Y1)
    accepted = np.zeros(ra1.size, dtype=bool)
    ring0 = np.zeros(ra1.size, dtype=bool)
    winner = np.full(ra1.size, -1, dtype=np.int64)
    shrink = 1.0 - g["eps_lo"] - 1e-9
    for i in range(ra1.size):
        x0, y0, x1, y1, c1 = int(sx[i]), int(sy[i]), X1[i], Y1[i], C1[i]
        best = [K_INF_HI, K_INF_HI, -1]

        def scan(k):
            qs, qe = cstart[k], cstart[k + 1]
            if qe == qs:
                return
            dx, dy = X2[qs:qe] - x1, Y2[qs:qe] - y1
            h = _hi(dy*dy + dx*dx*C2[qs:qe]*c1)
            if self_match:
                h = np.where(order[qs:qe] == i, K_INF_HI, h)
            for t in range(h.size):
                ht = int(h[t])
                better = ht < best[0]
                best[1] = min(best[1], max(best[0], ht))
                best[0] = min(best[0], ht)
                if better:
                    best[2] = qs + t

        key = x0*ndec + y0
        scan(key)                                                        # ring 0
        ex_lo, ey_lo = g["ox"] + x0*wx, g["oy"] + y0*wy
        dxc, dyc = (ex_lo + 0.5*wx) - x1, (ey_lo + 0.5*wy) - y1
        cdist = np.sqrt(dxc*dxc*g["ccen_row"][y0]*c1 + dyc*dyc)
        flag, done = False, False
        dec = stop_decision(_hi_floor(best[0]), _hi_ceil(best[0]), 0.4*g["cs_rad"] - 1.1*cdist, g)
        if dec == 0:
            done = True
        elif dec == 2:
            flag, done = True, True
        if not done:
            gl = max(x1 - ex_lo - g["fzx"], 0.0); gr = max((ex_lo + wx) - x1 - g["fzx"], 0.0)
            gd = max(y1 - ey_lo - g["fzy"], 0.0); gu = max((ey_lo + wy) - y1 - g["fzy"], 0.0)
            GX = [(gl + wx)*(gl + wx)*c1, gl*gl*c1, 0.0, gr*gr*c1, (gr + wx)*(gr + wx)*c1]
            GY = [(gd + wy)*(gd + wy)*shrink, gd*gd*shrink, 0.0, gu*gu*shrink, (gu + wy)*(gu + wy)*shrink]
            CM = []
            for k in range(5):
                yy = y0 + k - 2
                CM.append(g["cmin_row"][yy]*shrink if 0 <= yy < ndec else -1.0)
            thr = _hi_ceil(best[0])*(1.0 + g["eps_hi"])*(1.0 + 1e-9)
            cm_lo = min(CM)
            outer = (not (min(min(GX[0], GX[4])*max(cm_lo, 0.0), min(GY[0], GY[4])) >= thr)) or cm_lo < 0.0
            cells = []
            for e in range(24):
                if e == 8 and not outer:
                    break
                bx, by = BX[e], BY[e]
                cxx = x0 + bx
                cm = CM[by + 2]
                lb = GX[bx + 2]*cm + GY[by + 2]
                if 0 <= cxx < nra and cm >= 0.0 and lb < thr:
                    cells.append(key + bx*ndec + by)
            for k in cells:
                scan(k)
            if stop_decision(_hi_floor(best[0]), _hi_ceil(best[0]), 2.4*g["cs_rad"] - 1.1*cdist, g) != 0:
                flag = True
        if best[1] < K_INF_HI and not (_hi_ceil(best[0])*(1.0 + g["eps_hi"]) < _hi_floor(best[1])*(1.0 - g["eps_lo"])):
            flag = True
        if best[2] < 0:
            flag = True
        if not flag:
            accepted[i] = True
            ring0[i] = done
            winner[i] = order[best[2]]
    return accepted, winner, ring0
