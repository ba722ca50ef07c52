/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference hot path.
 *
 * Nothing under vif_b200/ may include, link, import or execute this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only
 * as the checker / CPU baseline. The product path is CUDA-only and fails loudly without it.
 *
 * What it restates (all citations relative to /root/reference/include/vif):
 *   astro/qxmatch.hpp:133-616   qxmatch(): argument conventions, unit conversion (:170-181),
 *                               distance proxy (:192-204), proxy<->arcsec (:206-222),
 *                               grid geometry (:232-267), bucket fill (:270-288),
 *                               forward ring search work1 (:293-341), reverse work2 (:343-380),
 *                               nth lowering (:383-387), brute force (:491-526),
 *                               final conversion (:609-613)
 *   astro/qxmatch.hpp:44-128    depth_cache ring table (quadrant + 3 mirrors, axis cells twice)
 *   astro/astro.hpp:25-39       angdistr / angdist
 *   astro/astro.hpp:524-552     field_area_hull (Heron fan over the hull, great-circle edges)
 *   math/convex_hull.hpp:108-150 monotone-chain hull (only ever fed 4 bbox corners here)
 *   math/reduce.hpp:46-53       mean() = (0.0 + a + b)/2
 *
 * Parity pin: the reference ships no test for this path, so this port is
 * pinned against the reference itself compiled here (oracle/_ref, oracle/ref_wrapper.cpp) by
 * tests/test_oracle_vs_ref.py and against the committed fixtures tests/golden/*.npz that were
 * minted from the reference by tests/golden/make_golden.py.
 *
 * Semantics are those of thread=1. `nthreads` only splits rows between
 * pthreads; results are independent of it (each output row/column is owned by one thread and
 * scans candidates in the reference's order).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (x86-64 gcc never contracts without -mfma; the flag
 * makes that explicit) -- see oracle/Makefile.
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NPOS UINT64_MAX
/* math/base.hpp:20: dpi is spelled with 18 digits and rounds to the double nearest pi. */
static const double DPI = 3.14159265358979323;

typedef struct {
    double rra0, rra1, rdec0, rdec1; /* padded bounds (qxmatch.hpp:260-263) */
    double cell_size, dra, ddec;     /* qxmatch.hpp:246-257 */
    uint64_t nra, ndec;              /* qxmatch.hpp:266-267 */
    uint64_t nc2;                    /* qxmatch.hpp:242 */
    double area;                     /* field_area_hull of cat-2 bbox corners, deg^2 */
} oracle_grid_t;

typedef struct {
    uint64_t n_eval_fwd, n_eval_rev; /* distance_proxy calls */
    uint64_t max_ring_fwd, max_ring_rev;
} oracle_info_t;

/* ---- angular distances (astro.hpp:25-39) ------------------------------------------------- */
static double angdistr(double ra1, double dec1, double ra2, double dec2) {
    double sra = sin(0.5*(ra2 - ra1));
    double sde = sin(0.5*(dec2 - dec1));
    return 2.0*asin(sqrt(sde*sde + sra*sra*cos(dec2)*cos(dec1)));
}

double oracle_angdist(double tra1, double tdec1, double tra2, double tdec2) {
    const double d2r = DPI/180.0;
    double ra1 = d2r*tra1, ra2 = d2r*tra2, dec1 = d2r*tdec1, dec2 = d2r*tdec2;
    double sra = sin(0.5*(ra2 - ra1));
    double sde = sin(0.5*(dec2 - dec1));
    return 3600.0*2.0*asin(sqrt(sde*sde + sra*sra*cos(dec2)*cos(dec1)))/d2r;
}

/* vector forms: astro.hpp:43-81 (element-wise; ra2/dec2 scalar when n2 == 1) */
void oracle_angdist_vec(const double* ra1, const double* dec1, uint64_t n, const double* ra2,
                        const double* dec2, uint64_t n2, double* out) {
    for (uint64_t i = 0; i < n; ++i)
        out[i] = oracle_angdist(ra1[i], dec1[i], ra2[n2 == 1 ? 0 : i], dec2[n2 == 1 ? 0 : i]);
}

/* angdist_less, astro.hpp:83-103: proxy < sqr(sin(rrad/2)), no asin */
void oracle_angdist_less(const double* tra1, const double* tdec1, uint64_t n, double tra2, double tdec2,
                         double radius, uint8_t* out) {
    const double d2r = DPI/180.0;
    const double rrad = d2r*radius/3600.0;
    const double sr = sin(rrad/2.0);
    const double crad = sr*sr;
    for (uint64_t i = 0; i < n; ++i) {
        double ra1 = d2r*tra1[i], ra2 = d2r*tra2;
        double dec1 = d2r*tdec1[i], dec2 = d2r*tdec2;
        double sra = sin(0.5*(ra2 - ra1));
        double sde = sin(0.5*(dec2 - dec1));
        out[i] = (sde*sde + sra*sra*cos(dec2)*cos(dec1) < crad) ? 1 : 0;
    }
}

/* qdist, qxmatch.hpp:684-717: n x n matrix, ret(j,i) for j > i, everything else 0 */
void oracle_qdist(const double* ra, const double* dec, uint64_t n, double* out) {
    const double d2r = DPI/180.0;
    double* dra = (double*)malloc(3*n*sizeof(double) + 8);
    double* ddec = dra + n;
    double* dcdec = ddec + n;
    for (uint64_t i = 0; i < n; ++i) { dra[i] = ra[i]*d2r; ddec[i] = dec[i]*d2r; dcdec[i] = cos(ddec[i]); }
    memset(out, 0, n*n*sizeof(double));
    for (uint64_t i = 0; i < n; ++i) {
        for (uint64_t j = i + 1; j < n; ++j) {
            double sra = sin(0.5*(dra[j] - dra[i]));
            double sde = sin(0.5*(ddec[j] - ddec[i]));
            double d = sde*sde + sra*sra*dcdec[j]*dcdec[i];
            out[j*n + i] = 3600.0*(180.0/DPI)*2*asin(sqrt(d));
        }
    }
    free(dra);
}

/* ---- hull of <= 4 points + spherical Heron area (convex_hull.hpp:108-150, astro.hpp:524-552) */
static double cross3(const double* x, const double* y, int i, int j, int k) {
    return (x[j] - x[i])*(y[k] - y[i]) - (y[j] - y[i])*(x[k] - x[i]);
}

double oracle_field_area_bbox(double r0, double r1, double d0, double d1) {
    /* corner order of qxmatch.hpp:243-244 */
    double x[4] = {r0, r0, r1, r1};
    double y[4] = {d0, d1, d0, d1};
    int ord[4] = {0, 1, 2, 3};
    /* lexicographic (x, then y) ordering; the comparator is a strict weak order so any sort
       gives the same sequence up to permutations of identical points */
    for (int a = 1; a < 4; ++a) {
        int v = ord[a], b = a;
        while (b > 0) {
            int u = ord[b-1];
            int less = (x[v] < x[u]) || (!(x[v] > x[u]) && y[v] < y[u]);
            if (!less) break;
            ord[b] = u; --b;
        }
        ord[b] = v;
    }
    int hull[8]; int k = 0;
    for (int i = 0; i < 4; ++i) {                       /* lower chain */
        while (k >= 2 && cross3(x, y, hull[k-2], hull[k-1], ord[i]) <= 0) --k;
        hull[k++] = ord[i];
    }
    int t = k + 1;
    for (int i = 2; i >= 0; --i) {                      /* upper chain, closes the polygon */
        while (k >= t && cross3(x, y, hull[k-2], hull[k-1], ord[i]) <= 0) --k;
        hull[k++] = ord[i];
    }
    const double d2r = DPI/180.0;
    double hx[8], hy[8];
    for (int i = 0; i < k; ++i) { hx[i] = x[hull[i]]*d2r; hy[i] = y[hull[i]]*d2r; }
    double area = 0;
    for (int i = 2; i < k; ++i) {                       /* fan from vertex 0, Heron's formula */
        double e1 = angdistr(hx[0],   hy[0],   hx[i-1], hy[i-1]);
        double e2 = angdistr(hx[i-1], hy[i-1], hx[i],   hy[i]);
        double e3 = angdistr(hx[i],   hy[i],   hx[0],   hy[0]);
        double p = 0.5*(e1 + e2 + e3);
        area += sqrt(p*(p-e1)*(p-e2)*(p-e3));
    }
    area /= d2r*d2r;
    return area;
}

/* ---- grid geometry (qxmatch.hpp:232-267) -------------------------------------------------- */
/* bbox = {ra_min, ra_max, dec_min, dec_max} of each catalog */
void oracle_grid(const double* bb1, const double* bb2, uint64_t n2, uint64_t nth, int linear,
                 oracle_grid_t* g) {
    double rra0  = bb1[0] < bb2[0] ? bb1[0] : bb2[0];
    double rra1  = bb1[1] > bb2[1] ? bb1[1] : bb2[1];
    double rdec0 = bb1[2] < bb2[2] ? bb1[2] : bb2[2];
    double rdec1 = bb1[3] > bb2[3] ? bb1[3] : bb2[3];

    const double overgrowth = 10.0;
    /* dpi*n2/nth/overgrowth evaluated left to right in double */
    uint64_t nc2 = (uint64_t)ceil(0.5*sqrt(DPI*(double)n2/(double)nth/overgrowth));
    double area = 0.0, cell_size;
    if (linear) {
        cell_size = sqrt((bb2[1] - bb2[0])*(bb2[3] - bb2[2]));
    } else {
        area = oracle_field_area_bbox(bb2[0], bb2[1], bb2[2], bb2[3]);
        double c = 3600.0*sqrt(area)/(double)nc2;
        cell_size = c > 1.0 ? c : 1.0;   /* std::max(c, 1.0) */
    }
    double dra, ddec;
    if (linear) {
        dra = cell_size; ddec = cell_size;
    } else {
        double mean_dec = (0.0 + bb2[2] + bb2[3])/2.0;   /* reduce.hpp:46-53 on {min,max} */
        dra = cell_size/fabs(cos(mean_dec*DPI/180.0))/3600.0;
        ddec = cell_size/3600.0;
    }
    rra0 -= dra; rra1 += dra; rdec0 -= ddec; rdec1 += ddec;
    g->rra0 = rra0; g->rra1 = rra1; g->rdec0 = rdec0; g->rdec1 = rdec1;
    g->cell_size = cell_size; g->dra = dra; g->ddec = ddec;
    g->nra  = (uint64_t)((rra1 - rra0)/dra);
    g->ndec = (uint64_t)((rdec1 - rdec0)/ddec);
    g->nc2 = nc2; g->area = area;
}

/* ---- ring table (qxmatch.hpp:44-128), grown lazily exactly like depth_cache --------------- */
typedef struct {
    int64_t *bx, *by; uint64_t n; double max_dist;
} ring_t;

typedef struct {
    ring_t* rings; uint64_t nrings, cap;
    uint8_t* visited; uint64_t nx, ny; double cs; uint64_t max_depth;
} ringtab_t;

static double sqr(double x) { return x*x; }

static void ring_grow(ringtab_t* t) {
    if (t->nrings == t->cap) {
        t->cap = t->cap ? 2*t->cap : 16;
        t->rings = (ring_t*)realloc(t->rings, t->cap*sizeof(ring_t));
    }
    ring_t* r = &t->rings[t->nrings++];
    uint64_t d = t->nrings;                    /* = index + 1 */
    r->max_dist = t->cs*((double)d + 0.5);
    uint64_t xm = d+1 < t->nx ? d+1 : t->nx;
    uint64_t ym = d+1 < t->ny ? d+1 : t->ny;
    uint64_t cap = 4*xm*ym + 4, q = 0;
    r->bx = (int64_t*)malloc(cap*sizeof(int64_t));
    r->by = (int64_t*)malloc(cap*sizeof(int64_t));
    for (uint64_t x = 0; x < xm; ++x)
    for (uint64_t y = 0; y < ym; ++y) {
        double c0 = sqr((double)x-0.5) + sqr((double)y-0.5);
        double c1 = sqr((double)x+0.5) + sqr((double)y-0.5);
        double c2 = sqr((double)x+0.5) + sqr((double)y+0.5);
        double c3 = sqr((double)x-0.5) + sqr((double)y+0.5);
        double m = c0; if (c1 < m) m = c1; if (c2 < m) m = c2; if (c3 < m) m = c3;
        double dist2 = sqr(t->cs)*m;
        if (dist2 <= sqr(r->max_dist) && !t->visited[x*t->ny + y]) {
            r->bx[q] = (int64_t)x; r->by[q] = (int64_t)y; ++q;
            t->visited[x*t->ny + y] = 1;
        }
    }
    /* mirrors in the order (-x,+y), (-x,-y), (+x,-y); axis cells therefore repeat */
    for (uint64_t i = 0; i < q; ++i) { r->bx[q+i]   = -r->bx[i]; r->by[q+i]   =  r->by[i]; }
    for (uint64_t i = 0; i < q; ++i) { r->bx[2*q+i] = -r->bx[i]; r->by[2*q+i] = -r->by[i]; }
    for (uint64_t i = 0; i < q; ++i) { r->bx[3*q+i] =  r->bx[i]; r->by[3*q+i] = -r->by[i]; }
    r->n = 4*q;
}

static void ringtab_init(ringtab_t* t, uint64_t nx, uint64_t ny, double cs) {
    memset(t, 0, sizeof(*t));
    t->nx = nx; t->ny = ny; t->cs = cs; t->max_depth = nx > ny ? nx : ny;
    t->visited = (uint8_t*)calloc(nx*ny ? nx*ny : 1, 1);
    t->cap = 16; t->rings = (ring_t*)malloc(t->cap*sizeof(ring_t));
    ring_t* r = &t->rings[0]; t->nrings = 1;
    r->bx = (int64_t*)malloc(sizeof(int64_t)); r->by = (int64_t*)malloc(sizeof(int64_t));
    r->bx[0] = 0; r->by[0] = 0; r->n = 1; r->max_dist = cs/2.0;
    if (nx*ny) t->visited[0] = 1;
}

static const ring_t* ringtab_get(ringtab_t* t, uint64_t i) {
    while (i >= t->nrings) ring_grow(t);
    return &t->rings[i];
}

static void ringtab_free(ringtab_t* t) {
    for (uint64_t i = 0; i < t->nrings; ++i) { free(t->rings[i].bx); free(t->rings[i].by); }
    free(t->rings); free(t->visited);
}

/* ring dump for tests: returns entry count of ring k */
uint64_t oracle_ring(uint64_t nx, uint64_t ny, double cs, uint64_t k,
                     int64_t* bx, int64_t* by, uint64_t cap, double* md) {
    ringtab_t t; ringtab_init(&t, nx, ny, cs);
    const ring_t* r = ringtab_get(&t, k);
    uint64_t n = r->n;
    for (uint64_t i = 0; i < n && i < cap; ++i) { bx[i] = r->bx[i]; by[i] = r->by[i]; }
    *md = r->max_dist;
    ringtab_free(&t);
    return n;
}

/* ---- shared state of one call ------------------------------------------------------------- */
typedef struct {
    const double *ra1, *dec1, *ra2, *dec2;       /* inputs (degrees, or linear units) */
    double *dra1, *ddec1, *dcdec1, *dra2, *ddec2, *dcdec2;
    uint64_t n1, n2, nth_alloc, nth;
    int self, no_mirror, linear;
    oracle_grid_t g;
    /* CSR buckets: ids in ascending order per cell (qxmatch.hpp:280-288 push_back order) */
    uint64_t *off1, *ids1, *off2, *ids2, *cx1, *cy1, *cx2, *cy2;
    uint64_t *id; double *d; uint64_t *rid; double *rd;
} call_t;

static double proxy(const call_t* c, uint64_t i, uint64_t j) {
    if (c->linear) {
        return sqr(c->ra2[j] - c->ra1[i]) + sqr(c->dec2[j] - c->dec1[i]);
    } else {
        double sra = sin(0.5*(c->dra2[j] - c->dra1[i]));
        double sde = sin(0.5*(c->ddec2[j] - c->ddec1[i]));
        return sde*sde + sra*sra*c->dcdec2[j]*c->dcdec1[i];
    }
}

static double true_to_proxy(const call_t* c, double dist) {
    if (c->linear) return sqr(dist);
    return sqr(sin(dist/(3600.0*(180.0/DPI)*2)));
}

static double proxy_to_true(const call_t* c, double p) {
    if (c->linear) return sqrt(p);
    return 3600.0*(180.0/DPI)*2*asin(sqrt(p));
}

static double true_distance(const call_t* c, double x1, double y1, double x2, double y2) {
    if (c->linear) return sqrt(sqr(x2 - x1) + sqr(y2 - y1));
    return oracle_angdist(x1, y1, x2, y2);
}

/* insertion into the sorted nth-list of row i (qxmatch.hpp:323-332): strict <, bubble with > */
static void insert(const call_t* c, uint64_t i, uint64_t j, double sd) {
    const uint64_t n1 = c->n1, nth = c->nth;
    if (sd < c->d[(nth-1)*n1 + i]) {
        c->id[(nth-1)*n1 + i] = j;
        c->d[(nth-1)*n1 + i] = sd;
        for (uint64_t k = nth-1; k-- > 0; ) {
            double *a = &c->d[k*n1 + i], *b = &c->d[(k+1)*n1 + i];
            if (!(*a > *b)) break;
            double td = *a; *a = *b; *b = td;
            uint64_t *ia = &c->id[k*n1 + i], *ib = &c->id[(k+1)*n1 + i];
            uint64_t ti = *ia; *ia = *ib; *ib = ti;
        }
    }
}

typedef struct { call_t* c; uint64_t b1, e1, b2, e2; uint64_t nev1, nev2, mr1, mr2; } job_t;

static void* bucket_job(void* arg) {
    job_t* jb = (job_t*)arg; call_t* c = jb->c;
    const oracle_grid_t* g = &c->g;
    ringtab_t tab; ringtab_init(&tab, g->nra, g->ndec, g->cell_size);
    const uint64_t n1 = c->n1;
    for (uint64_t i = jb->b1; i < jb->e1; ++i) {                 /* work1 */
        int64_t x0 = (int64_t)c->cx1[i], y0 = (int64_t)c->cy1[i];
        double cell_dist = true_distance(c, c->ra1[i], c->dec1[i],
            g->rra0 + ((double)x0+0.5)*g->dra, g->rdec0 + ((double)y0+0.5)*g->ddec);
        uint64_t d = 0; double reached;
        do {
            const ring_t* r = ringtab_get(&tab, d);
            for (uint64_t b = 0; b < r->n; ++b) {
                int64_t x = x0 + r->bx[b], y = y0 + r->by[b];
                if (x < 0 || (uint64_t)x >= g->nra || y < 0 || (uint64_t)y >= g->ndec) continue;
                uint64_t cell = (uint64_t)x*g->ndec + (uint64_t)y;
                for (uint64_t q = c->off2[cell]; q < c->off2[cell+1]; ++q) {
                    uint64_t j = c->ids2[q];
                    if (c->self && i == j) continue;
                    ++jb->nev1;
                    insert(c, i, j, proxy(c, i, j));
                }
            }
            /* re-fetch: the table may have been reallocated by a grow */
            r = ringtab_get(&tab, d);
            double t = r->max_dist - 1.1*cell_dist - 0.1*g->cell_size;
            reached = true_to_proxy(c, t > 0.0 ? t : 0.0);      /* std::max(0.0, t) */
            if (d > jb->mr1) jb->mr1 = d;
            ++d;
        } while (c->d[(c->nth-1)*n1 + i] > reached && d < tab.max_depth);
    }
    if (!c->self && !c->no_mirror) {
        for (uint64_t j = jb->b2; j < jb->e2; ++j) {             /* work2 */
            int64_t x0 = (int64_t)c->cx2[j], y0 = (int64_t)c->cy2[j];
            double cell_dist = true_distance(c, c->ra2[j], c->dec2[j],
                g->rra0 + ((double)x0+0.5)*g->dra, g->rdec0 + ((double)y0+0.5)*g->ddec);
            uint64_t d = 0; double reached;
            do {
                const ring_t* r = ringtab_get(&tab, d);
                for (uint64_t b = 0; b < r->n; ++b) {
                    int64_t x = x0 + r->bx[b], y = y0 + r->by[b];
                    if (x < 0 || (uint64_t)x >= g->nra || y < 0 || (uint64_t)y >= g->ndec) continue;
                    uint64_t cell = (uint64_t)x*g->ndec + (uint64_t)y;
                    for (uint64_t q = c->off1[cell]; q < c->off1[cell+1]; ++q) {
                        uint64_t i = c->ids1[q];
                        ++jb->nev2;
                        double sd = proxy(c, i, j);
                        if (sd < c->rd[j]) { c->rd[j] = sd; c->rid[j] = i; }
                    }
                }
                r = ringtab_get(&tab, d);
                double t = r->max_dist - 1.1*cell_dist - 0.1*g->cell_size;
                reached = true_to_proxy(c, t > 0.0 ? t : 0.0);
                if (d > jb->mr2) jb->mr2 = d;
                ++d;
            } while (c->rd[j] > reached && d < tab.max_depth);
        }
    }
    ringtab_free(&tab);
    return NULL;
}

static void* brute_job(void* arg) {
    job_t* jb = (job_t*)arg; call_t* c = jb->c;
    /* forward: rows [b1,e1), ascending j, strict < (qxmatch.hpp:519-523, 498-507) */
    for (uint64_t i = jb->b1; i < jb->e1; ++i)
        for (uint64_t j = 0; j < c->n2; ++j) {
            if (c->self && i == j) continue;
            ++jb->nev1;
            insert(c, i, j, proxy(c, i, j));
        }
    /* reverse: columns [b2,e2), ascending i, strict < == thread=1 rule (qxmatch.hpp:510-513) */
    if (!c->no_mirror)
        for (uint64_t j = jb->b2; j < jb->e2; ++j)
            for (uint64_t i = 0; i < c->n1; ++i) {
                if (c->self && i == j) continue;
                ++jb->nev2;
                double sd = proxy(c, i, j);
                if (sd < c->rd[j]) { c->rd[j] = sd; c->rid[j] = i; }
            }
    return NULL;
}

static void build_csr(const double* ra, const double* dec, uint64_t n, const oracle_grid_t* g,
                      uint64_t** cx, uint64_t** cy, uint64_t** off, uint64_t** ids) {
    uint64_t ncell = g->nra*g->ndec;
    *cx = (uint64_t*)malloc((n ? n : 1)*sizeof(uint64_t));
    *cy = (uint64_t*)malloc((n ? n : 1)*sizeof(uint64_t));
    *off = (uint64_t*)calloc(ncell + 2, sizeof(uint64_t));
    *ids = (uint64_t*)malloc((n ? n : 1)*sizeof(uint64_t));
    for (uint64_t i = 0; i < n; ++i) {
        /* qxmatch.hpp:278-279 on the raw (degree) inputs */
        (*cx)[i] = (uint64_t)floor((ra[i] - g->rra0)/g->dra);
        (*cy)[i] = (uint64_t)floor((dec[i] - g->rdec0)/g->ddec);
        ++(*off)[(*cx)[i]*g->ndec + (*cy)[i] + 2];
    }
    for (uint64_t k = 2; k < ncell + 2; ++k) (*off)[k] += (*off)[k-1];
    /* off[c+1] is now the insertion cursor of cell c; ascending i keeps push_back order */
    for (uint64_t i = 0; i < n; ++i) (*ids)[(*off)[(*cx)[i]*g->ndec + (*cy)[i] + 1]++] = i;
}

/*
 * Entry point. Output buffers are caller-allocated: id,d = nth*n1 with nth = max(nth_in,1)
 * (plane-major: id[k*n1+i], qxmatch.hpp:157), rid,rd = n2 (ignored when no_mirror).
 * Returns 0, or 1 for non-finite input (the reference prints and exits, qxmatch.hpp:146-149).
 */
int oracle_qxmatch(const double* ra1, const double* dec1, uint64_t n1,
                   const double* ra2, const double* dec2, uint64_t n2,
                   uint64_t nth_in, uint64_t nthreads, int self, int brute_force,
                   int no_mirror, int linear,
                   uint64_t* id, double* d, uint64_t* rid, double* rd,
                   oracle_grid_t* grid_out, oracle_info_t* info) {
    for (uint64_t i = 0; i < n1; ++i) if (!isfinite(ra1[i]) || !isfinite(dec1[i])) return 1;
    for (uint64_t j = 0; j < n2; ++j) if (!isfinite(ra2[j]) || !isfinite(dec2[j])) return 1;

    call_t c; memset(&c, 0, sizeof(c));
    c.ra1 = ra1; c.dec1 = dec1; c.ra2 = ra2; c.dec2 = dec2; c.n1 = n1; c.n2 = n2;
    c.nth_alloc = nth_in < 1 ? 1 : nth_in; c.nth = c.nth_alloc;
    c.self = self; c.no_mirror = no_mirror; c.linear = linear;
    c.id = id; c.d = d; c.rid = rid; c.rd = rd;
    for (uint64_t k = 0; k < c.nth_alloc*n1; ++k) { id[k] = NPOS; d[k] = INFINITY; }
    if (!no_mirror) for (uint64_t j = 0; j < n2; ++j) { rid[j] = NPOS; rd[j] = INFINITY; }
    if (info) memset(info, 0, sizeof(*info));
    if (grid_out) memset(grid_out, 0, sizeof(*grid_out));
    if (n1 == 0 || n2 == 0) return 0;                       /* qxmatch.hpp:165-167: d stays inf */

    if (!linear) {
        const double d2r = DPI/180.0;
        c.dra1 = (double*)malloc(n1*sizeof(double)); c.ddec1 = (double*)malloc(n1*sizeof(double));
        c.dcdec1 = (double*)malloc(n1*sizeof(double));
        c.dra2 = (double*)malloc(n2*sizeof(double)); c.ddec2 = (double*)malloc(n2*sizeof(double));
        c.dcdec2 = (double*)malloc(n2*sizeof(double));
        for (uint64_t i = 0; i < n1; ++i) {
            c.dra1[i] = ra1[i]*d2r; c.ddec1[i] = dec1[i]*d2r; c.dcdec1[i] = cos(c.ddec1[i]);
        }
        for (uint64_t j = 0; j < n2; ++j) {
            c.dra2[j] = ra2[j]*d2r; c.ddec2[j] = dec2[j]*d2r; c.dcdec2[j] = cos(c.ddec2[j]);
        }
    }

    if (nthreads < 1) nthreads = 1;
    job_t* jobs = (job_t*)calloc(nthreads, sizeof(job_t));
    pthread_t* th = (pthread_t*)malloc(nthreads*sizeof(pthread_t));

    if (!brute_force) {
        double bb1[4] = {ra1[0], ra1[0], dec1[0], dec1[0]}, bb2[4] = {ra2[0], ra2[0], dec2[0], dec2[0]};
        for (uint64_t i = 1; i < n1; ++i) {
            if (ra1[i] < bb1[0]) bb1[0] = ra1[i];
            if (ra1[i] > bb1[1]) bb1[1] = ra1[i];
            if (dec1[i] < bb1[2]) bb1[2] = dec1[i];
            if (dec1[i] > bb1[3]) bb1[3] = dec1[i];
        }
        for (uint64_t j = 1; j < n2; ++j) {
            if (ra2[j] < bb2[0]) bb2[0] = ra2[j];
            if (ra2[j] > bb2[1]) bb2[1] = ra2[j];
            if (dec2[j] < bb2[2]) bb2[2] = dec2[j];
            if (dec2[j] > bb2[3]) bb2[3] = dec2[j];
        }
        oracle_grid(bb1, bb2, n2, c.nth_alloc, linear, &c.g);
        if (grid_out) *grid_out = c.g;
        build_csr(ra1, dec1, n1, &c.g, &c.cx1, &c.cy1, &c.off1, &c.ids1);
        build_csr(ra2, dec2, n2, &c.g, &c.cx2, &c.cy2, &c.off2, &c.ids2);
    }
    if (n2 < c.nth) c.nth = n2;       /* qxmatch.hpp:383-387: after allocation, planes >= n2 stay npos */
    /* (the brute-force lambda captures nth before any lowering, qxmatch.hpp:491; with n2 < nth it
       simply never fills the last planes, which is the same outcome) */
    if (brute_force) c.nth = c.nth_alloc;

    for (uint64_t t = 0; t < nthreads; ++t) {
        jobs[t].c = &c;
        jobs[t].b1 = n1*t/nthreads; jobs[t].e1 = n1*(t+1)/nthreads;
        jobs[t].b2 = n2*t/nthreads; jobs[t].e2 = n2*(t+1)/nthreads;
    }
    void* (*fn)(void*) = brute_force ? brute_job : bucket_job;
    if (nthreads == 1) fn(&jobs[0]);
    else {
        for (uint64_t t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, fn, &jobs[t]);
        for (uint64_t t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    }
    if (info) for (uint64_t t = 0; t < nthreads; ++t) {
        info->n_eval_fwd += jobs[t].nev1; info->n_eval_rev += jobs[t].nev2;
        if (jobs[t].mr1 > info->max_ring_fwd) info->max_ring_fwd = jobs[t].mr1;
        if (jobs[t].mr2 > info->max_ring_rev) info->max_ring_rev = jobs[t].mr2;
    }

    for (uint64_t k = 0; k < c.nth_alloc*n1; ++k) d[k] = proxy_to_true(&c, d[k]);   /* inf -> NaN */
    if (!no_mirror) for (uint64_t j = 0; j < n2; ++j) rd[j] = proxy_to_true(&c, rd[j]);

    free(jobs); free(th);
    free(c.dra1); free(c.ddec1); free(c.dcdec1); free(c.dra2); free(c.ddec2); free(c.dcdec2);
    free(c.off1); free(c.ids1); free(c.off2); free(c.ids2);
    free(c.cx1); free(c.cy1); free(c.cx2); free(c.cy2);
    return 0;
}
