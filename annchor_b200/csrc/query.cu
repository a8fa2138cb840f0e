// query.cu -- Annchor.query(): out-of-sample queries against a fitted index
// (annchor/annchor.py:643-683, annchor/query_functions.py:10-212).
//
// The reference builds, for the nq x nx rectangle of (point, query) pairs, the same objects as fit():
// query-to-anchor distances, the locality candidate set, bounds / dad features, the stored regression's
// clipped prediction and error label, per-query thresholds, guarantee_nmin, probabilities, ONE refine
// round of the n_refine most probable pairs, and a per-query top-nn over the computed pairs.  Here the
// rectangle is processed in batches of queries whose RefineApprox row block (4 B), label (1 B) and
// probability level (2 B) per pair are resident in HBM: unlike fit()'s N^2/2 pairs the rectangle of one
// batch is bounded (<= 2^30 pairs), so the stages are plain streaming kernels at HBM rate:
//   fill      : 128 B of point-major anchor distances in, 5 B out per pair
//   select    : per-query k-th smallest by 4-pass radix select over the row (thresh, guarantee_nmin)
//   level     : probability level per not-computed candidate + global level histogram
//   emit      : pairs above the cut level, ties at the cut by the smallest tie_key (as in fit())
//   top-nn    : per query, nn extract-min rounds ordered by (distance, point id)
// The arithmetic is that of the fit() sweeps (float32, same fused multiply-add chain), ties are
// resolved by the same rules, so oracle/devmode.py restates it exactly.
#include <algorithm>
#include <cmath>

#include "index_obj.cuh"

namespace annb {

int pair_dists_f64(annb_ctx *c, const annb_dataset *ds, int metric, const int32_t *I, const int32_t *J, int64_t n,
                   double *out);
int pair_dists_f32_perm(annb_ctx *c, const annb_dataset *ds, int metric, const int32_t *I, const int32_t *J,
                        const int32_t *perm, int64_t n, float *out);

constexpr uint8_t QL_NONE = 0xff;      // not a candidate
constexpr uint8_t QL_COMPUTED = 0x80;  // flag: exact value present
constexpr uint16_t QV_NONE = 0xffff;   // no probability level (computed / not a candidate)

struct QMeta {
    uint64_t amask;  // the query's `locality` nearest anchors (query_functions.py:29)
    int32_t cA;      // its closest anchor (np_argmin(QD, 1), query_functions.py:94)
    int32_t pad;
};

// per query: float32 copy of its anchor distances, nearest-anchor mask (ties to the lower index), argmin
__global__ void query_meta_kernel(const double *__restrict__ QD64 /* (nq, na) */, int64_t nq, int na, int locality,
                                  float *__restrict__ QD32 /* (nq, 64) */, QMeta *__restrict__ qm)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < nq; j += (int64_t)gridDim.x * blockDim.x) {
        QMeta m;
        m.amask = 0;
        m.cA = 0;
        m.pad = 0;
        double best = INFINITY;
        for (int a = 0; a < na; ++a) {
            const double v = QD64[j * na + a];
            QD32[j * kMaxAnchors + a] = (float)v;
            if (v < best) {
                best = v;
                m.cA = a;
            }
        }
        for (int a = na; a < kMaxAnchors; ++a) QD32[j * kMaxAnchors + a] = 0.0f;
        for (int l = 0; l < locality && l < na; ++l) {
            double bv = INFINITY;
            int ba = -1;
            for (int a = 0; a < na; ++a) {
                if ((m.amask >> a) & 1ull) continue;
                const double v = QD64[j * na + a];
                if (ba < 0 || v < bv) {
                    bv = v;
                    ba = a;
                }
            }
            m.amask |= 1ull << ba;
        }
        qm[j] = m;
    }
}

// doubled-edge bin / label and the clipped prediction, exactly as the fit() sweeps compute them (sweep.cuh)
__device__ __forceinline__ float q_predict(const Model &M, float lb, float ub, float s2, int &label)
{
    int b = 0, l = 0;
#pragma unroll
    for (int k = 1; k < MAX_BINS; ++k) {
        b += s2 > M.e2[k];
        l += s2 >= M.e2[k];
    }
    label = l;
    const float y = fmaf(lb, M.c0[b], fmaf(ub, M.c1[b], fmaf(s2, 0.5f * M.c2[b], M.ic[b])));
    return fminf(fmaxf(y, lb), ub);
}

// RefineApprox / label of every (point i, query j) pair of the batch (query_functions.py:40-68,183-204)
__global__ void __launch_bounds__(256)
query_fill_kernel(View V, const __grid_constant__ Model M, const float *__restrict__ QD32, const QMeta *__restrict__ qm,
                  int64_t nq, int loc_thresh, float *__restrict__ RA, uint8_t *__restrict__ lab)
{
    __shared__ float s_q[kMaxAnchors];
    const int64_t j = blockIdx.y;
    if (threadIdx.x < kMaxAnchors) s_q[threadIdx.x] = QD32[j * kMaxAnchors + threadIdx.x];
    __syncthreads();
    const QMeta q = qm[j];
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < V.n; i += (int64_t)gridDim.x * blockDim.x) {
        const PointMeta pm = V.meta[i];
        float v = 0.0f;
        uint8_t l = QL_NONE;
        if (__popcll(pm.amask & q.amask) >= loc_thresh) {
            const float *di = V.Dpm + i * V.dpitch;
            float lb = 0.0f, ub = INFINITY;
            for (int a0 = 0; a0 < V.na; a0 += 4) {
                const float4 x = __ldg(reinterpret_cast<const float4 *>(di + a0));
                const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (a0 + t < V.na) {
                        lb = fmaxf(lb, fabsf(xs[t] - s_q[a0 + t]));
                        ub = fminf(ub, xs[t] + s_q[a0 + t]);
                    }
            }
            const float s2 = di[q.cA] + s_q[pm.cA];  // 2 * dad (query_functions.py:96-98)
            int label;
            v = q_predict(M, lb, ub, s2, label);
            l = (uint8_t)label | (pm.slot >= 0 ? QL_COMPUTED : 0);  // np.isin(IJs[:, 0], ann.A), :62
        }
        RA[j * V.n + i] = v;
        lab[j * V.n + i] = l;
    }
}

__device__ __forceinline__ uint32_t f2key(float f)  // order-preserving float -> uint
{
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// per query: the kq[j]-th smallest (0-based) value among its candidates (mode 0) or its not-computed
// candidates (mode 1); 4-pass radix select over the row.  kq[j] < 0 or too few values: +inf.
// Also counts the computed candidates of the row (mode 0).
__global__ void __launch_bounds__(256)
query_select_kernel(const float *__restrict__ RA, const uint8_t *__restrict__ lab, int64_t nx, const int32_t *__restrict__ kq,
                    int mode, float *__restrict__ out, int32_t *__restrict__ n_computed)
{
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_k, s_ok;
    __shared__ int s_comp;
    const int64_t j = blockIdx.x;
    const float *row = RA + j * nx;
    const uint8_t *lrow = lab + j * nx;
    const int want = kq[j];
    if (threadIdx.x == 0) {
        s_prefix = 0;
        s_k = want < 0 ? 0u : (uint32_t)want;
        s_ok = want >= 0;
        s_comp = 0;
    }
    __syncthreads();
    for (int shift = 24; shift >= 0; shift -= 8) {
        hist[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        int comp = 0;
        for (int64_t i = threadIdx.x; i < nx; i += blockDim.x) {
            const uint8_t l = lrow[i];
            if (l == QL_NONE) continue;
            if (shift == 24 && (l & QL_COMPUTED)) ++comp;
            if (mode == 1 && (l & QL_COMPUTED)) continue;
            const uint32_t key = f2key(row[i]);
            if (shift < 24 && (key >> (shift + 8)) != prefix) continue;
            atomicAdd(&hist[(key >> shift) & 255u], 1u);
        }
        if (shift == 24 && n_computed) atomicAdd(&s_comp, comp);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t k = s_k;
            int d = 0;
            for (; d < 256; ++d) {
                if (k < hist[d]) break;
                k -= hist[d];
            }
            if (d == 256) s_ok = 0;  // fewer than kq + 1 values
            s_k = k;
            s_prefix = (prefix << 8) | (uint32_t)(d & 255);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[j] = s_ok ? key2f(s_prefix) : INFINITY;
        if (n_computed) n_computed[j] = s_comp;
    }
}

// guarantee_nmin (utils.py:606-621 as called at query_functions.py:144-150): not-computed candidates
// strictly below the row's (n_todo + 1)-th smallest not-computed value are forced (RefineApprox = -1)
__global__ void query_force_kernel(float *__restrict__ RA, const uint8_t *__restrict__ lab, int64_t nx, int64_t nq,
                                   const float *__restrict__ kth, unsigned long long *__restrict__ n_forced)
{
    unsigned long long mine = 0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nq * nx; p += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t l = lab[p];
        if (l == QL_NONE || (l & QL_COMPUTED)) continue;
        const float t = kth[p / nx];
        if (t < INFINITY && RA[p] < t) {
            RA[p] = -1.0f;
            ++mine;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(n_forced, mine);
}

// probability level of every not-computed candidate: p = thresh[j] - RefineApprox (query_functions.py:152),
// level = rank of searchsorted(errs[label], p, 'left') / len among all such fractions (get_probs)
__global__ void __launch_bounds__(256)
query_level_kernel(const float *__restrict__ RA, const uint8_t *__restrict__ lab, int64_t nx, int64_t nq,
                   const float *__restrict__ thresh, const __grid_constant__ Model M, const float *__restrict__ errs,
                   const uint16_t *__restrict__ ranktab, int nlevels, uint16_t *__restrict__ lvl,
                   unsigned long long *__restrict__ hist)
{
    extern __shared__ uint32_t s_hist[];
    for (int k = threadIdx.x; k < nlevels; k += blockDim.x) s_hist[k] = 0;
    __syncthreads();
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nq * nx; p += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t l = lab[p];
        uint16_t out = QV_NONE;
        if (l != QL_NONE && !(l & QL_COMPUTED)) {
            const int label = l & 7;
            const float pr = thresh[p / nx] - RA[p];
            const float *er = errs + M.eoff[label];
            int lo = 0, hi = M.eoff[label + 1] - M.eoff[label];
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (er[mid] < pr) lo = mid + 1;
                else hi = mid;
            }
            out = ranktab[M.eoff[label] + label + lo];
            atomicAdd(&s_hist[out], 1u);
        }
        lvl[p] = out;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < nlevels; k += blockDim.x)
        if (s_hist[k]) atomicAdd(&hist[k], (unsigned long long)s_hist[k]);
}

// tie-break key of a (point, query) pair; the query id is offset by nx so that keys are those of the
// pair (i, nx + j) in the data set "X followed by Q"
__device__ __forceinline__ uint64_t q_tie_key(int64_t p, int64_t nx, int64_t q0, uint64_t salt)
{
    return tie_key((uint32_t)(p % nx), (uint32_t)(nx + q0 + p / nx), salt);
}

// 16-bit digit histogram of the tie keys of the pairs at `level` whose higher digits equal `prefix`
__global__ void query_tie_hist_kernel(const uint16_t *__restrict__ lvl, int64_t npairs, int64_t nx, int64_t q0,
                                      int level, uint64_t salt, uint64_t prefix, int shift, uint32_t *__restrict__ hist)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
        if (lvl[p] != level) continue;
        const uint64_t k = q_tie_key(p, nx, q0, salt);
        if (shift < 48 && (k >> (shift + 16)) != prefix) continue;
        atomicAdd(&hist[(k >> shift) & 0xffff], 1u);
    }
}

// selected = level above the cut, or at the cut with tie key <= thr
__global__ void query_emit_kernel(const uint16_t *__restrict__ lvl, int64_t npairs, int64_t nx, int64_t q0, int cut,
                                  uint64_t salt, uint64_t thr, int32_t *__restrict__ I, int32_t *__restrict__ J,
                                  int64_t *__restrict__ pos, int64_t cap, unsigned long long *__restrict__ cnt)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < npairs; p += (int64_t)gridDim.x * blockDim.x) {
        const uint16_t l = lvl[p];
        if (l == QV_NONE || (int)l < cut) continue;
        if ((int)l == cut && q_tie_key(p, nx, q0, salt) > thr) continue;
        const unsigned long long s = atomicAdd(cnt, 1ull);
        if ((int64_t)s < cap) {
            I[s] = (int32_t)(p % nx);
            J[s] = (int32_t)(nx + q0 + p / nx);
            pos[s] = p;
        }
    }
}

__global__ void query_store_kernel(const int64_t *__restrict__ pos, const float *__restrict__ d, int64_t m,
                                   float *__restrict__ RA, uint8_t *__restrict__ lab)
{
    for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < m; s += (int64_t)gridDim.x * blockDim.x) {
        RA[pos[s]] = d[s];
        lab[pos[s]] |= QL_COMPUTED;
    }
}

// get_nn(nq, nn + 1, ...) (query_functions.py:208, utils.py:383-429): per query the nn smallest computed
// values ordered by (value, point id); if a row has fewer computed pairs the remaining slots take the
// not-computed candidates in order of their RefineApprox, which is also what is emitted
__global__ void __launch_bounds__(256)
query_topk_kernel(const float *__restrict__ RA, const uint8_t *__restrict__ lab, int64_t nx, int nn,
                  int64_t *__restrict__ ngi, double *__restrict__ ngd)
{
    __shared__ float s_v[8];
    __shared__ int s_i[8];
    __shared__ float s_pv;
    __shared__ int s_pi, s_phase;
    const int64_t j = blockIdx.x;
    const float *row = RA + j * nx;
    const uint8_t *lrow = lab + j * nx;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        s_pv = -INFINITY;
        s_pi = -1;
        s_phase = 0;  // 0: computed pairs, 1: not-computed candidates
    }
    __syncthreads();
    for (int r = 0; r < nn; ++r) {
        for (;;) {
            const float pv = s_pv;
            const int pi = s_pi, phase = s_phase;
            float bv = INFINITY;
            int bi = INT32_MAX;
            for (int64_t i = threadIdx.x; i < nx; i += blockDim.x) {
                const uint8_t l = lrow[i];
                if (l == QL_NONE || ((l & QL_COMPUTED) != 0) != (phase == 0)) continue;
                const float v = row[i];
                const bool after = v > pv || (v == pv && (int)i > pi);
                if (after && (v < bv || (v == bv && (int)i < bi))) {
                    bv = v;
                    bi = (int)i;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float yv = __shfl_xor_sync(0xffffffffu, bv, o);
                const int yi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (yv < bv || (yv == bv && yi < bi)) {
                    bv = yv;
                    bi = yi;
                }
            }
            if (lane == 0) {
                s_v[w] = bv;
                s_i[w] = bi;
            }
            __syncthreads();
            bool again = false;
            if (threadIdx.x == 0) {
                for (int q = 1; q < 8; ++q)
                    if (s_v[q] < bv || (s_v[q] == bv && s_i[q] < bi)) {
                        bv = s_v[q];
                        bi = s_i[q];
                    }
                if (bi == INT32_MAX && phase == 0) {  // computed pairs exhausted: go on with predictions
                    s_phase = 1;
                    s_pv = -INFINITY;
                    s_pi = -1;
                } else {
                    s_pv = bi == INT32_MAX ? INFINITY : bv;
                    s_pi = bi;
                    ngi[j * nn + r] = bi == INT32_MAX ? -1 : bi;
                    ngd[j * nn + r] = bi == INT32_MAX ? INFINITY : (double)bv;
                }
            }
            __syncthreads();
            again = s_phase != phase;
            if (!again) break;
        }
    }
}

static int qgrid(const annb_ctx *c, int64_t n)
{
    const int64_t g = (n + 255) / 256, cap = (int64_t)c->num_sms * 16;
    return (int)std::max<int64_t>(1, std::min(g, cap));
}

}  // namespace annb

using namespace annb;

// get_exact_query_ijs(f, X, Z, IJ) (annchor/utils.py:180-245): out[p] = metric(X[i_p], Z[j_p]) where
// `both` holds the nx items of X followed by the items of Z
ANNB_API int annb_pair_dists_query(annb_ctx *c, const annb_dataset *both, int metric, int64_t nx, const int64_t *ij,
                                   int64_t n, double *out)
{
    ANNB_REQUIRE(c && both && (n == 0 || (ij && out)), ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(nx > 0 && nx <= both->n, ANNB_EINVAL, "nx outside the combined data set");
    std::vector<int64_t> t((size_t)2 * n);
    for (int64_t p = 0; p < n; ++p) {
        ANNB_REQUIRE(ij[2 * p] >= 0 && ij[2 * p] < nx && ij[2 * p + 1] >= 0 && ij[2 * p + 1] < both->n - nx, ANNB_EINVAL,
                     "query pair index out of range");
        t[2 * p] = ij[2 * p];
        t[2 * p + 1] = nx + ij[2 * p + 1];
    }
    return annb_pair_dists(c, both, metric, t.data(), n, out);
}

ANNB_API int annb_index_query(annb_index *ix, const annb_dataset *both, int64_t nq, int64_t nn, double p_work,
                              int64_t *ngi, double *ngd, int64_t *n_evals)
{
    TraceScope _ts("annb_index_query");
    ANNB_REQUIRE(ix && both && ngi && ngd, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(ix->have_anchors && ix->have_model && ix->nlevels > 0, ANNB_ESTATE,
                 "query needs a fitted index (anchors, regression and error model)");
    ANNB_REQUIRE(!ix->A_host.empty(), ANNB_ESTATE, "query needs anchors that are points of X");
    annb_ctx *c = ix->ctx;
    const int64_t nx = ix->n;
    const int na = ix->na;
    ANNB_REQUIRE(nq > 0 && both->n == nx + nq, ANNB_EINVAL, "the combined data set must hold X followed by the %lld queries",
                 (long long)nq);
    ANNB_REQUIRE(nn >= 1 && nn < MAX_LIST && nn <= nx, ANNB_ERANGE, "nn=%lld outside [1, %d]", (long long)nn, MAX_LIST - 1);
    ANNB_REQUIRE(p_work > 0, ANNB_EINVAL, "p_work must be positive");
    ANNB_CUDA(cudaSetDevice(c->device));
    if (n_evals) *n_evals = 0;
    const View V = ix->view();
    DevBuf bI, bJ, bQD64, bQD32, bQM, bRA, bLab, bLvl, bKq, bKth, bTh, bComp, bHist, bTie, bPos, bD, bOutI, bOutD;
    int rc = ANNB_OK;
    auto cleanup = [&]() {
        for (DevBuf *b : {&bI, &bJ, &bQD64, &bQD32, &bQM, &bRA, &bLab, &bLvl, &bKq, &bKth, &bTh, &bComp, &bHist, &bTie, &bPos,
                          &bD, &bOutI, &bOutD})
            b->release();
    };
#define Q_TRY(expr)          \
    do {                     \
        rc = (expr);         \
        if (rc != ANNB_OK) { \
            cleanup();       \
            return rc;       \
        }                    \
    } while (0)
#define Q_CUDA(expr)                                                              \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) {                                                  \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            cleanup();                                                            \
            return ANNB_ECUDA;                                                    \
        }                                                                         \
    } while (0)
    // batches of queries whose rectangle stays below 2^30 pairs (7 B per pair resident)
    const int64_t qb = std::max<int64_t>(1, std::min<int64_t>(nq, ((int64_t)1 << 30) / nx));
    const int nlev = ix->nlevels;
    int64_t evals = 0;
    std::vector<uint64_t> hist(nlev);
    std::vector<uint32_t> th(65536);
    for (int64_t q0 = 0; q0 < nq; q0 += qb) {
        const int64_t mq = std::min<int64_t>(qb, nq - q0);
        const int64_t np = mq * nx;
        // 1. query -> anchor distances (query_functions.py:10-15): na exact evaluations per query
        Q_TRY(bI.ensure((size_t)std::max<int64_t>(mq * na, 1) * 4));
        Q_TRY(bJ.ensure((size_t)std::max<int64_t>(mq * na, 1) * 4));
        {
            std::vector<int32_t> hi((size_t)mq * na), hj((size_t)mq * na);
            for (int64_t j = 0; j < mq; ++j)
                for (int a = 0; a < na; ++a) {
                    hi[(size_t)j * na + a] = ix->A_host[a];
                    hj[(size_t)j * na + a] = (int32_t)(nx + q0 + j);
                }
            Q_CUDA(cudaMemcpyAsync(bI.p, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice, c->stream));
            Q_CUDA(cudaMemcpyAsync(bJ.p, hj.data(), hj.size() * 4, cudaMemcpyHostToDevice, c->stream));
            Q_CUDA(cudaStreamSynchronize(c->stream));
        }
        Q_TRY(bQD64.ensure((size_t)mq * na * 8));
        Q_TRY(pair_dists_f64(c, both, ix->metric, bI.as<int32_t>(), bJ.as<int32_t>(), mq * na, bQD64.as<double>()));
        Q_TRY(bQD32.ensure((size_t)mq * kMaxAnchors * 4));
        Q_TRY(bQM.ensure((size_t)mq * sizeof(QMeta)));
        query_meta_kernel<<<qgrid(c, mq), 256, 0, c->stream>>>(bQD64.as<double>(), mq, na, ix->P.locality,
                                                               bQD32.as<float>(), bQM.as<QMeta>());
        ++g_launches;
        evals += mq * na;
        // 2. features, prediction, labels of the whole rectangle
        Q_TRY(bRA.ensure((size_t)np * 4));
        Q_TRY(bLab.ensure((size_t)np));
        Q_TRY(bLvl.ensure((size_t)np * 2));
        {
            dim3 grid((unsigned)std::min<int64_t>((nx + 255) / 256, 64), (unsigned)mq);
            query_fill_kernel<<<grid, 256, 0, c->stream>>>(V, ix->model, bQD32.as<float>(), bQM.as<QMeta>(), mq,
                                                          ix->P.loc_thresh, bRA.as<float>(), bLab.as<uint8_t>());
            ++g_launches;
        }
        // 3. thresh = (nn + 1)-th smallest RefineApprox per query (query_functions.py:142)
        Q_TRY(bKq.ensure((size_t)mq * 4));
        Q_TRY(bKth.ensure((size_t)mq * 4));
        Q_TRY(bTh.ensure((size_t)mq * 4));
        Q_TRY(bComp.ensure((size_t)mq * 4));
        std::vector<int32_t> kq((size_t)mq, (int32_t)nn);
        Q_CUDA(cudaMemcpyAsync(bKq.p, kq.data(), (size_t)mq * 4, cudaMemcpyHostToDevice, c->stream));
        query_select_kernel<<<(unsigned)mq, 256, 0, c->stream>>>(bRA.as<float>(), bLab.as<uint8_t>(), nx, bKq.as<int32_t>(), 0,
                                                               bTh.as<float>(), bComp.as<int32_t>());
        ++g_launches;
        // 4. guarantee_nmin with nmin = 3 nn / 2 (query_functions.py:144-150)
        std::vector<int32_t> comp((size_t)mq);
        Q_CUDA(cudaMemcpyAsync(comp.data(), bComp.p, (size_t)mq * 4, cudaMemcpyDeviceToHost, c->stream));
        Q_CUDA(cudaStreamSynchronize(c->stream));
        const int64_t nmin = 3 * nn / 2;
        bool any = false;
        for (int64_t j = 0; j < mq; ++j) {
            const int64_t todo = nmin - comp[j];
            kq[j] = todo > 0 ? (int32_t)todo : -1;
            any |= todo > 0;
        }
        Q_TRY(bHist.ensure((size_t)std::max(nlev, 16) * 8 + 64));
        unsigned long long *counters = bHist.as<unsigned long long>() + nlev;
        Q_CUDA(cudaMemsetAsync(bHist.p, 0, (size_t)nlev * 8 + 64, c->stream));
        if (any) {
            Q_CUDA(cudaMemcpyAsync(bKq.p, kq.data(), (size_t)mq * 4, cudaMemcpyHostToDevice, c->stream));
            query_select_kernel<<<(unsigned)mq, 256, 0, c->stream>>>(bRA.as<float>(), bLab.as<uint8_t>(), nx,
                                                                   bKq.as<int32_t>(), 1, bKth.as<float>(), nullptr);
            query_force_kernel<<<qgrid(c, np), 256, 0, c->stream>>>(bRA.as<float>(), bLab.as<uint8_t>(), nx, mq,
                                                                   bKth.as<float>(), counters);
            g_launches += 2;
        }
        // 5. probability levels + global histogram
        Q_CUDA(cudaFuncSetAttribute(query_level_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, nlev * 4));
        query_level_kernel<<<c->num_sms * 4, 256, (size_t)nlev * 4, c->stream>>>(
            bRA.as<float>(), bLab.as<uint8_t>(), nx, mq, bTh.as<float>(), ix->model, ix->errs_dev.as<float>(),
            ix->rank_dev.as<uint16_t>(), nlev, bLvl.as<uint16_t>(), bHist.as<unsigned long long>());
        ++g_launches;
        Q_CUDA(cudaMemcpyAsync(hist.data(), bHist.p, (size_t)nlev * 8, cudaMemcpyDeviceToHost, c->stream));
        Q_CUDA(cudaStreamSynchronize(c->stream));
        int64_t n_nc = 0;
        for (int l = 0; l < nlev; ++l) n_nc += (int64_t)hist[l];
        // n_refine = int(p_work * nq * nx - na * nq) + 1 (query_functions.py:170-174), this batch's share
        int64_t n_refine = (int64_t)(p_work * (double)mq * (double)nx - (double)na * (double)mq) + 1;
        n_refine = std::max<int64_t>(0, std::min(n_refine, n_nc));
        // 6. cut: level first, then the smallest tie keys at the cut level
        int cut = nlev;
        int64_t above = 0, t = 0;
        for (int l = nlev - 1; l >= 0 && n_refine > 0; --l) {
            if (above + (int64_t)hist[l] >= n_refine) {
                cut = l;
                t = n_refine - above;
                break;
            }
            above += (int64_t)hist[l];
        }
        ix->n_selects += 1;
        const uint64_t salt = mix64(0x9E3779B97F4A7C15ull * (uint64_t)ix->n_selects);
        uint64_t thr = ~0ull;
        if (cut < nlev && t < (int64_t)hist[cut]) {
            Q_TRY(bTie.ensure(65536 * 4));
            uint64_t prefix = 0;
            int64_t tt = t;
            for (int shift = 48; shift >= 0; shift -= 16) {
                Q_CUDA(cudaMemsetAsync(bTie.p, 0, 65536 * 4, c->stream));
                query_tie_hist_kernel<<<qgrid(c, np), 256, 0, c->stream>>>(bLvl.as<uint16_t>(), np, nx, q0, cut, salt, prefix,
                                                                          shift, bTie.as<uint32_t>());
                ++g_launches;
                Q_CUDA(cudaMemcpyAsync(th.data(), bTie.p, 65536 * 4, cudaMemcpyDeviceToHost, c->stream));
                Q_CUDA(cudaStreamSynchronize(c->stream));
                int d = 0;
                for (; d < 65536; ++d) {
                    if (tt <= (int64_t)th[d]) break;
                    tt -= th[d];
                }
                if (d == 65536) {
                    set_error("query: tie selection ran past the histogram");
                    cleanup();
                    return ANNB_ESTATE;
                }
                prefix = (prefix << 16) | (uint64_t)d;
            }
            thr = prefix;
        }
        // 7. emit, evaluate, store
        if (n_refine > 0) {
            Q_TRY(bI.ensure((size_t)n_refine * 4));
            Q_TRY(bJ.ensure((size_t)n_refine * 4));
            Q_TRY(bPos.ensure((size_t)n_refine * 8));
            Q_TRY(bD.ensure((size_t)n_refine * 4));
            Q_CUDA(cudaMemsetAsync(counters, 0, 64, c->stream));
            query_emit_kernel<<<qgrid(c, np), 256, 0, c->stream>>>(bLvl.as<uint16_t>(), np, nx, q0, cut, salt, thr,
                                                                  bI.as<int32_t>(), bJ.as<int32_t>(), bPos.as<int64_t>(),
                                                                  n_refine, counters + 1);
            ++g_launches;
            unsigned long long got = 0;
            Q_CUDA(cudaMemcpyAsync(&got, counters + 1, 8, cudaMemcpyDeviceToHost, c->stream));
            Q_CUDA(cudaStreamSynchronize(c->stream));
            if ((int64_t)got != n_refine) {
                set_error("query: selection produced %llu pairs for a target of %lld", got, (long long)n_refine);
                cleanup();
                return ANNB_ESTATE;
            }
            Q_TRY(pair_dists_f32_perm(c, both, ix->metric, bI.as<int32_t>(), bJ.as<int32_t>(), nullptr, n_refine,
                                      bD.as<float>()));
            query_store_kernel<<<qgrid(c, n_refine), 256, 0, c->stream>>>(bPos.as<int64_t>(), bD.as<float>(), n_refine,
                                                                         bRA.as<float>(), bLab.as<uint8_t>());
            ++g_launches;
            evals += n_refine;
        }
        // 8. per-query top-nn
        Q_TRY(bOutI.ensure((size_t)mq * nn * 8));
        Q_TRY(bOutD.ensure((size_t)mq * nn * 8));
        query_topk_kernel<<<(unsigned)mq, 256, 0, c->stream>>>(bRA.as<float>(), bLab.as<uint8_t>(), nx, (int)nn,
                                                             bOutI.as<int64_t>(), bOutD.as<double>());
        ++g_launches;
        Q_CUDA(cudaMemcpyAsync(ngi + q0 * nn, bOutI.p, (size_t)mq * nn * 8, cudaMemcpyDeviceToHost, c->stream));
        Q_CUDA(cudaMemcpyAsync(ngd + q0 * nn, bOutD.p, (size_t)mq * nn * 8, cudaMemcpyDeviceToHost, c->stream));
        Q_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (n_evals) *n_evals = evals;
    cleanup();
#undef Q_TRY
#undef Q_CUDA
    return ANNB_OK;
}
