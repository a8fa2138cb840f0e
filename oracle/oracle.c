/*
 * oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C (OpenMP) CPU restatement of the numeric leaves of gchq/annchor's
 * `Annchor.fit()` hot path.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product (annchor_b200/) never does.
 *
 * Every function cites the reference file:line whose arithmetic it restates
 * (paths relative to the upstream repository root).  The metrics whose
 * arithmetic lives in un-vendored wheels (python-Levenshtein==0.27.1 ->
 * RapidFuzz==3.13.0, pynndescent==0.5.13 kantorovich) are restated from their
 * published definitions and pinned against the reference's own known-answer
 * tests and bundled exact 100-NN fixtures (see tests/golden/, tests/test_oracle.py).
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared -fPIC)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* Metrics (annchor/distances.py:8-20, annchor/utils.py:62-86)          */
/* ------------------------------------------------------------------ */

/* Unit-cost edit distance (ins = del = sub = 1) over code points: the
 * published definition of Levenshtein.distance as called at
 * annchor/distances.py:16-20.  Two-row Wagner-Fischer DP. */
static int64_t lev_dp(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb,
                      int32_t *row)
{
    if (la == 0) return lb;
    if (lb == 0) return la;
    for (int64_t j = 0; j <= lb; ++j) row[j] = (int32_t)j;
    for (int64_t i = 1; i <= la; ++i) {
        int32_t diag = row[0];
        row[0] = (int32_t)i;
        const uint8_t ca = a[i - 1];
        for (int64_t j = 1; j <= lb; ++j) {
            int32_t up = row[j];
            int32_t sub = diag + (ca != b[j - 1]);
            int32_t best = up + 1 < row[j - 1] + 1 ? up + 1 : row[j - 1] + 1;
            row[j] = sub < best ? sub : best;
            diag = up;
        }
    }
    return row[lb];
}

API int64_t orc_lev(const uint8_t *a, int64_t la, const uint8_t *b, int64_t lb)
{
    int32_t *row = (int32_t *)malloc((size_t)(lb + 1) * sizeof(int32_t));
    int64_t r = lev_dp(a, la, b, lb, row);
    free(row);
    return r;
}

/* get_exact_ijs (annchor/utils.py:144-150) specialised to levenshtein over a
 * packed byte corpus: chars + offsets[n+1]. */
API void orc_lev_pairs(const uint8_t *chars, const int64_t *offs, const int64_t *ij,
                       int64_t n, double *out)
{
#pragma omp parallel
    {
        int64_t cap = 0;
        int32_t *row = NULL;
#pragma omp for schedule(dynamic, 64)
        for (int64_t p = 0; p < n; ++p) {
            int64_t i = ij[2 * p], j = ij[2 * p + 1];
            int64_t la = offs[i + 1] - offs[i], lb = offs[j + 1] - offs[j];
            if (lb + 1 > cap) {
                cap = lb + 1;
                row = (int32_t *)realloc(row, (size_t)cap * sizeof(int32_t));
            }
            out[p] = (double)lev_dp(chars + offs[i], la, chars + offs[j], lb, row);
        }
        free(row);
    }
}

/* euclidean = np.linalg.norm(x - y) (annchor/distances.py:8-13); the result has
 * the precision of the input dtype and is stored into a float64 array
 * (annchor/utils.py:146-149). */
API void orc_euclid_pairs_f32(const float *X, int64_t d, const int64_t *ij, int64_t n,
                              double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p) {
        const float *x = X + ij[2 * p] * d, *y = X + ij[2 * p + 1] * d;
        double s = 0.0;
        for (int64_t k = 0; k < d; ++k) {
            float t = x[k] - y[k];
            s += (double)t * (double)t;
        }
        out[p] = (double)(float)sqrt(s);
    }
}

API void orc_euclid_pairs_f64(const double *X, int64_t d, const int64_t *ij, int64_t n,
                              double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p) {
        const double *x = X + ij[2 * p] * d, *y = X + ij[2 * p + 1] * d;
        double s = 0.0;
        for (int64_t k = 0; k < d; ++k) {
            double t = x[k] - y[k];
            s += t * t;
        }
        out[p] = sqrt(s);
    }
}

/* cosine = scipy.spatial.distance.cosine (annchor/utils.py:14,67):
 * 1 - u.v / sqrt(u.u * v.v), clipped to [0, 2]. */
API void orc_cosine_pairs_f32(const float *X, int64_t d, const int64_t *ij, int64_t n,
                              double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p) {
        const float *x = X + ij[2 * p] * d, *y = X + ij[2 * p + 1] * d;
        double uv = 0, uu = 0, vv = 0;
        for (int64_t k = 0; k < d; ++k) {
            uv += (double)x[k] * y[k];
            uu += (double)x[k] * x[k];
            vv += (double)y[k] * y[k];
        }
        double r = 1.0 - uv / sqrt(uu * vv);
        out[p] = r < 0 ? 0 : (r > 2 ? 2 : r);
    }
}

API void orc_cosine_pairs_f64(const double *X, int64_t d, const int64_t *ij, int64_t n,
                              double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p) {
        const double *x = X + ij[2 * p] * d, *y = X + ij[2 * p + 1] * d;
        double uv = 0, uu = 0, vv = 0;
        for (int64_t k = 0; k < d; ++k) {
            uv += x[k] * y[k];
            uu += x[k] * x[k];
            vv += y[k] * y[k];
        }
        double r = 1.0 - uv / sqrt(uu * vv);
        out[p] = r < 0 ? 0 : (r > 2 ? 2 : r);
    }
}

/* 'wasserstein' (annchor/utils.py:75-86) for the 1-D ground cost |a-b| on the
 * bin index: kantorovich(x, y, cost) normalises each histogram to unit mass and
 * solves exact OT; with cost |a-b| the optimum is sum_b |CDF_x(b) - CDF_y(b)|. */
API void orc_w1_pairs_f64(const double *H, int64_t nb, const int64_t *ij, int64_t n,
                          double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < n; ++p) {
        const double *x = H + ij[2 * p] * nb, *y = H + ij[2 * p + 1] * nb;
        double sx = 0, sy = 0;
        for (int64_t k = 0; k < nb; ++k) {
            sx += x[k];
            sy += y[k];
        }
        double cx = 0, cy = 0, w = 0;
        for (int64_t k = 0; k < nb; ++k) {
            cx += x[k] / sx;
            cy += y[k] / sy;
            w += fabs(cx - cy);
        }
        out[p] = w;
    }
}

/* 'wasserstein' with a general cost matrix (annchor/utils.py:75-86): kantorovich(x, y, cost=M) of
 * pynndescent 0.5.13 (absent from /root/reference; its published algorithm: drop the empty bins,
 * normalise both histograms to unit mass, solve the transportation problem exactly by network simplex and
 * return the optimal cost).  The optimum is unique, so any exact solver restates it: this one is the
 * textbook successive-shortest-path method (Dijkstra with node potentials on the dense bipartite residual
 * graph), scalar float64.  Pinned against the reference's bundled digits 100-NN distances
 * (tests/golden/digits.npz) and against an LP solver in tests/test_oracle.py. */
#define OT_MAXB 256
static double ot_pair(const double *x, const double *y, int64_t nb, const double *C)
{
    int si[OT_MAXB], tj[OT_MAXB], prev_s[OT_MAXB], prev_t[OT_MAXB];
    char vs[OT_MAXB], vt[OT_MAXB];
    double ra[OT_MAXB], rb[OT_MAXB], ps[OT_MAXB], pt[OT_MAXB], ds[OT_MAXB], dt[OT_MAXB];
    static __thread double f[OT_MAXB * OT_MAXB];
    const double EPS = 1e-13, TINY = 1e-15;
    double sx = 0, sy = 0;
    for (int64_t k = 0; k < nb; ++k) {
        sx += x[k];
        sy += y[k];
    }
    int m = 0, n = 0;
    for (int64_t k = 0; k < nb; ++k) {
        if (x[k] > 0) {
            si[m] = (int)k;
            ra[m++] = x[k] / sx;
        }
        if (y[k] > 0) {
            tj[n] = (int)k;
            rb[n++] = y[k] / sy;
        }
    }
    if (m == 0 || n == 0) return 0.0;
    for (int q = 0; q < m * n; ++q) f[q] = 0.0;
    for (int k = 0; k < m; ++k) ps[k] = 0.0;
    for (int l = 0; l < n; ++l) {
        double mn = INFINITY;
        for (int k = 0; k < m; ++k)
            if (C[si[k] * nb + tj[l]] < mn) mn = C[si[k] * nb + tj[l]];
        pt[l] = mn;
    }
    for (int round = 0; round < 8 * (m + n) + 32; ++round) {
        double rem = 0;
        for (int k = 0; k < m; ++k)
            if (ra[k] > EPS) rem += ra[k];
        if (!(rem > EPS)) break;
        for (int k = 0; k < m; ++k) {
            ds[k] = ra[k] > EPS ? 0.0 : INFINITY;
            prev_s[k] = -1;
            vs[k] = 0;
        }
        for (int l = 0; l < n; ++l) {
            dt[l] = INFINITY;
            prev_t[l] = -1;
            vt[l] = 0;
        }
        int target = -1;
        double dtar = INFINITY;
        for (;;) {
            double bv = INFINITY;
            int bid = -1;
            for (int k = 0; k < m; ++k)
                if (!vs[k] && ds[k] < bv) {
                    bv = ds[k];
                    bid = k;
                }
            for (int l = 0; l < n; ++l)
                if (!vt[l] && dt[l] < bv) {
                    bv = dt[l];
                    bid = OT_MAXB + l;
                }
            if (bid < 0) break;
            if (bid >= OT_MAXB) {
                const int l = bid - OT_MAXB;
                vt[l] = 1;
                if (rb[l] > EPS) {
                    target = l;
                    dtar = bv;
                    break;
                }
                for (int k = 0; k < m; ++k)
                    if (!vs[k] && f[k * n + l] > TINY) {
                        double rc = pt[l] - ps[k] - C[si[k] * nb + tj[l]];
                        if (rc < 0) rc = 0;
                        if (bv + rc < ds[k]) {
                            ds[k] = bv + rc;
                            prev_s[k] = l;
                        }
                    }
            } else {
                const int k = bid;
                vs[k] = 1;
                for (int l = 0; l < n; ++l)
                    if (!vt[l]) {
                        double rc = C[si[k] * nb + tj[l]] + ps[k] - pt[l];
                        if (rc < 0) rc = 0;
                        if (bv + rc < dt[l]) {
                            dt[l] = bv + rc;
                            prev_t[l] = k;
                        }
                    }
            }
        }
        if (target < 0) break;
        for (int k = 0; k < m; ++k) ps[k] += vs[k] ? ds[k] : dtar;
        for (int l = 0; l < n; ++l) pt[l] += vt[l] ? dt[l] : dtar;
        double amount = rb[target];
        int l = target, k;
        for (;;) {
            k = prev_t[l];
            if (prev_s[k] < 0) break;
            const int l2 = prev_s[k];
            if (f[k * n + l2] < amount) amount = f[k * n + l2];
            l = l2;
        }
        if (ra[k] < amount) amount = ra[k];
        const int root = k;
        l = target;
        for (;;) {
            k = prev_t[l];
            f[k * n + l] += amount;
            if (prev_s[k] < 0) break;
            const int l2 = prev_s[k];
            const double r = f[k * n + l2] - amount;
            f[k * n + l2] = r > TINY ? r : 0.0;
            l = l2;
        }
        ra[root] -= amount;
        rb[target] -= amount;
    }
    double cost = 0;
    for (int k = 0; k < m; ++k)
        for (int l = 0; l < n; ++l) cost += f[k * n + l] * C[si[k] * nb + tj[l]];
    return cost;
}

API int orc_ot_pairs_f64(const double *H, int64_t nb, const double *C, const int64_t *ij, int64_t n, double *out)
{
    if (nb > OT_MAXB) return -1;
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t p = 0; p < n; ++p)
        out[p] = ij[2 * p] == ij[2 * p + 1] ? 0.0 : ot_pair(H + ij[2 * p] * nb, H + ij[2 * p + 1] * nb, nb, C);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Bound / feature assembly (annchor/utils.py:274-301, 355-380)        */
/* D is (nx, na) row-major here (the reference passes a transposed view) */
/* ------------------------------------------------------------------ */

API void orc_bounds_ijs(const int64_t *ij, int64_t n, const double *D, int64_t na,
                        double *bounds /* (n,2) */)
{
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        const double *di = D + ij[2 * k] * na, *dj = D + ij[2 * k + 1] * na;
        double lo = -INFINITY, hi = INFINITY;
        for (int64_t a = 0; a < na; ++a) {
            double df = fabs(di[a] - dj[a]), sm = di[a] + dj[a];
            if (df > lo) lo = df;
            if (sm < hi) hi = sm;
        }
        bounds[2 * k] = lo;
        bounds[2 * k + 1] = hi;
    }
}

/* cA = argmin over anchors (first minimum, np.argmin); dad = (D[i,cA[j]] +
 * D[j,cA[i]]) / 2   (annchor/utils.py:375-380). */
API void orc_dad_ijs(const int64_t *ij, int64_t n, const double *D, int64_t nx, int64_t na,
                     double *dad)
{
    int64_t *cA = (int64_t *)malloc((size_t)nx * sizeof(int64_t));
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nx; ++i) {
        const double *di = D + i * na;
        int64_t best = 0;
        for (int64_t a = 1; a < na; ++a)
            if (di[a] < di[best]) best = a;
        cA[i] = best;
    }
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        int64_t i = ij[2 * k], j = ij[2 * k + 1];
        dad[k] = (D[i * na + cA[j]] + D[j * na + cA[i]]) / 2;
    }
    free(cA);
}

/* get_bounds_alt merge-join over sorted id lists (annchor/utils.py:304-323) and
 * update_bounds (annchor/utils.py:326-352).  CSR: ptr[nx+1], ids sorted per row. */
API void orc_update_bounds(const int64_t *ij, int64_t n, const int64_t *ptr,
                           const int64_t *ids, const double *ds, double *bounds /* (n,2) */)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t k = 0; k < n; ++k) {
        int64_t i = ij[2 * k], j = ij[2 * k + 1];
        int64_t p = ptr[i], pe = ptr[i + 1], q = ptr[j], qe = ptr[j + 1];
        double ub = INFINITY, lb = 0;
        while (p < pe && q < qe) {
            if (ids[p] < ids[q]) ++p;
            else if (ids[p] > ids[q]) ++q;
            else {
                double a = ds[p] + ds[q], b = fabs(ds[p] - ds[q]);
                if (a < ub) ub = a;
                if (b > lb) lb = b;
                ++p;
                ++q;
            }
        }
        bounds[2 * k] = lb;
        bounds[2 * k + 1] = ub;
    }
}

/* ------------------------------------------------------------------ */
/* Row selection helpers                                                */
/* ------------------------------------------------------------------ */



/* k-th smallest (0-based) by Hoare quickselect; buf is permuted. */
static double kth_smallest(double *a, int64_t n, int64_t k)
{
    int64_t lo = 0, hi = n - 1;
    while (lo < hi) {
        double pv = a[lo + ((hi - lo) >> 1)];
        int64_t i = lo, j = hi;
        while (i <= j) {
            while (a[i] < pv) ++i;
            while (a[j] > pv) --j;
            if (i <= j) {
                double t = a[i];
                a[i] = a[j];
                a[j] = t;
                ++i;
                --j;
            }
        }
        if (k <= j) hi = j;
        else if (k >= i) lo = i;
        else break;
    }
    return a[k];
}

/* thresh[i] = np.partition(RA[I[i]], nn)[nn]   (annchor/annchor.py:399-404).
 * CSR rows: ptr[nx+1], idx = pair indices touching row i. */
API void orc_row_kth(const double *RA, const int64_t *ptr, const int64_t *idx, int64_t nx,
                     int64_t k, double *out)
{
#pragma omp parallel
    {
        double *buf = NULL;
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < nx; ++i) {
            int64_t m = ptr[i + 1] - ptr[i];
            if (m > cap) {
                cap = m;
                buf = (double *)realloc(buf, (size_t)cap * sizeof(double));
            }
            for (int64_t t = 0; t < m; ++t) buf[t] = RA[idx[ptr[i] + t]];
            out[i] = kth_smallest(buf, m, k < m ? k : m - 1);
        }
        free(buf);
    }
}

/* get_probs (annchor/utils.py:581-589): searchsorted(errs[label], p, 'left') / len */
API void orc_probs(const double *p, const int64_t *label, int64_t n, const double *errs,
                   const int64_t *eptr /* nlabels+1 */, double *prob)
{
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        const double *e = errs + eptr[label[k]];
        int64_t len = eptr[label[k] + 1] - eptr[label[k]];
        int64_t lo = 0, hi = len;
        while (lo < hi) {
            int64_t mid = (lo + hi) >> 1;
            if (e[mid] < p[k]) lo = mid + 1;
            else hi = mid;
        }
        prob[k] = (double)lo / (double)len;
    }
}

typedef struct { double d; int64_t pair; } dp_t;
static int cmp_dp(const void *a, const void *b)
{
    const dp_t *x = (const dp_t *)a, *y = (const dp_t *)b;
    if (x->d != y->d) return (x->d > y->d) - (x->d < y->d);
    return (x->pair > y->pair) - (x->pair < y->pair);
}

/* get_nn (annchor/utils.py:383-429).  Ties are resolved by pair index (the
 * reference's unstable argsort leaves tie order unspecified). */
API void orc_get_nn(int64_t nx, int64_t nn, const double *RA, const int64_t *ij,
                    const int64_t *ptr, const int64_t *idx, const uint8_t *ncm, int64_t *ngi,
                    double *ngd)
{
#pragma omp parallel
    {
        dp_t *buf = NULL;
        int64_t cap = 0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t i = 0; i < nx; ++i) {
            int64_t m = ptr[i + 1] - ptr[i];
            if (m > cap) {
                cap = m;
                buf = (dp_t *)realloc(buf, (size_t)cap * sizeof(dp_t));
            }
            double mx = -INFINITY;
            for (int64_t t = 0; t < m; ++t) {
                double v = RA[idx[ptr[i] + t]];
                if (v > mx) mx = v;
            }
            for (int64_t t = 0; t < m; ++t) {
                int64_t pr = idx[ptr[i] + t];
                buf[t].d = RA[pr] + (ncm[pr] ? mx : 0.0);
                buf[t].pair = pr;
            }
            qsort(buf, (size_t)m, sizeof(dp_t), cmp_dp);
            for (int64_t t = 0; t < nn - 1; ++t) {
                int64_t pr = buf[t].pair;
                ngd[i * (nn - 1) + t] = RA[pr];
                ngi[i * (nn - 1) + t] = ij[2 * pr] == i ? ij[2 * pr + 1] : ij[2 * pr];
            }
        }
        free(buf);
    }
}

/* ------------------------------------------------------------------ */
/* numba's np.random inside @njit (annchor/utils.py:555-557, 572):      */
/* MT19937 init_genrand + Fisher-Yates from the end with bit-mask       */
/* rejection randint (numba/cpython/randomimpl.py get_next_int,         */
/* _randrange_impl 'np' flavour, do_shuffle_impl).                      */
/* ------------------------------------------------------------------ */

typedef struct { uint32_t mt[624]; int idx; } mt_t;

API void orc_mt_seed(mt_t *s, uint32_t seed)
{
    s->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->idx = 624;
}

static uint32_t mt_next(mt_t *s)
{
    if (s->idx >= 624) {
        uint32_t *mt = s->mt;
        for (int k = 0; k < 624; ++k) {
            uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
            mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        s->idx = 0;
    }
    uint32_t y = s->mt[s->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

static int64_t mt_randint(mt_t *s, int64_t n) /* uniform in [0, n) */
{
    if (n == 1) return 0;
    int nbits = 64 - __builtin_clzll((uint64_t)(n - 1));
    for (;;) {
        int64_t r;
        if (nbits <= 32) {
            r = (int64_t)(mt_next(s) & (0xffffffffu >> (32 - nbits)));
        } else {
            uint64_t hi = mt_next(s) & (0xffffffffu >> (64 - nbits));
            uint64_t lo = mt_next(s);
            r = (int64_t)((hi << 32) | lo);
        }
        if (r < n) return r;
    }
}

API size_t orc_mt_state_size(void) { return sizeof(mt_t); }

/* In-place np.random.shuffle as compiled by numba. */
API void orc_numba_shuffle(mt_t *s, int64_t *x, int64_t n)
{
    for (int64_t i = n - 1; i > 0; --i) {
        int64_t j = mt_randint(s, i + 1);
        int64_t t = x[i];
        x[i] = x[j];
        x[j] = t;
    }
}

/* ------------------------------------------------------------------ */
/* Device-arithmetic mode (oracle/devmode.py).  The CUDA sweeps compute  */
/* the same quantities as annchor/utils.py:274-301,355-380 and           */
/* annchor/regressors.py:71-103 + annchor/annchor.py:359-363, but in      */
/* float32 from a float32 copy of D and with a fused multiply-add chain.  */
/* These restate that arithmetic operation for operation (IEEE float32,  */
/* fmaf correctly rounded) so stage outputs can be compared for equality. */
/* ------------------------------------------------------------------ */

/* lb = max_a |D[i,a] - D[j,a]|, ub = min_a (D[i,a] + D[j,a]), s2 = D[i,cA[j]] + D[j,cA[i]] (= 2 dad) */
API void orc_f32_features(const int64_t *ij, int64_t n, const float *D, int64_t na,
                          const int32_t *cA, float *lb, float *ub, float *s2)
{
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        const int64_t i = ij[2 * k], j = ij[2 * k + 1];
        const float *di = D + i * na, *dj = D + j * na;
        float lo = 0.0f, hi = INFINITY;
        for (int64_t a = 0; a < na; ++a) {
            const float df = fabsf(di[a] - dj[a]), sm = di[a] + dj[a];
            if (df > lo) lo = df;
            if (sm < hi) hi = sm;
        }
        lb[k] = lo;
        ub[k] = hi;
        s2[k] = di[cA[j]] + dj[cA[i]];
    }
}

/* bin = #{interior doubled edges strictly below s2} ((lo, hi], regressors.py:85-87);
 * label = #{interior doubled edges <= s2} (closed, later bins win, error_predictors.py:63-66);
 * pred = clip(fma(lb, c0, fma(ub, c1, fma(s2, c2/2, icpt))), lb, ub).
 * e2[8]: e2[k] = 2*edge[k] for k = 1..nb-1, +inf otherwise; cf (8,4) = (c0, c1, c2/2, icpt). */
API void orc_f32_predict(const float *lb, const float *ub, const float *s2, int64_t n,
                         const float *e2, const float *cf, float *pred, int8_t *bin,
                         int8_t *label)
{
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < n; ++k) {
        int b = 0, l = 0;
        for (int q = 1; q < 8; ++q) {
            b += s2[k] > e2[q];
            l += s2[k] >= e2[q];
        }
        const float *c = cf + 4 * b;
        const float y = fmaf(lb[k], c[0], fmaf(ub[k], c[1], fmaf(s2[k], c[2], c[3])));
        pred[k] = fminf(fmaxf(y, lb[k]), ub[k]);
        if (bin) bin[k] = (int8_t)b;
        if (label) label[k] = (int8_t)l;
    }
}

/* update_bounds (annchor/utils.py:304-352) in float32 over CSR lists sorted by id */
API void orc_f32_update_bounds(const int64_t *ij, int64_t n, const int64_t *ptr,
                               const int64_t *ids, const float *ds, float *lbo, float *ubo)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t k = 0; k < n; ++k) {
        int64_t i = ij[2 * k], j = ij[2 * k + 1];
        int64_t p = ptr[i], pe = ptr[i + 1], q = ptr[j], qe = ptr[j + 1];
        float ub = INFINITY, lb = 0.0f;
        while (p < pe && q < qe) {
            if (ids[p] < ids[q]) ++p;
            else if (ids[p] > ids[q]) ++q;
            else {
                const float a = ds[p] + ds[q], b = fabsf(ds[p] - ds[q]);
                if (a < ub) ub = a;
                if (b > lb) lb = b;
                ++p;
                ++q;
            }
        }
        lbo[k] = lb;
        ubo[k] = ub;
    }
}
