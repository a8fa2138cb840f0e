// stage_ops.cu -- reference-shaped stage operators on EXPLICIT pair lists, float64.
// Drop-in replacements for the numba leaves of annchor/utils.py and the predict() methods of
// annchor/regressors.py / annchor/error_predictors.py, taking and returning the reference's
// own array layouts (host pointers; the entry points do the H2D / D2H copies).
// All arithmetic is float64 with the reference's operation order, so results are bit-identical
// for the max/min/abs/compare stages and within 1 ulp for the 3-term regression dot product.
#include "common.cuh"

namespace annb {

int split_ij(annb_ctx *c, const int64_t *ij_dev, int64_t n, int32_t *I, int32_t *J);

// --- get_bounds_njit_ijs (annchor/utils.py:274-301): one warp per pair, lanes over anchors ---
__global__ void __launch_bounds__(256)
bounds_ijs_kernel(const int64_t *__restrict__ ij, int64_t n, const double *__restrict__ D, int na,
                  double *__restrict__ bounds)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n; p += nwarps) {
        const double *di = D + ij[2 * p] * na, *dj = D + ij[2 * p + 1] * na;
        double lo = -INFINITY, hi = INFINITY;
        for (int a = lane; a < na; a += 32) {
            const double x = di[a], y = dj[a];
            lo = fmax(lo, fabs(x - y));
            hi = fmin(hi, x + y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmax(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmin(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) {
            bounds[2 * p] = lo;
            bounds[2 * p + 1] = hi;
        }
    }
}

// --- get_dad_ijs (annchor/utils.py:355-380) ---
__global__ void closest_anchor_kernel(const double *__restrict__ D, int64_t nx, int na,
                                      int32_t *__restrict__ cA)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nx;
         i += (int64_t)gridDim.x * blockDim.x) {
        const double *di = D + i * na;
        int best = 0;
        for (int a = 1; a < na; ++a)
            if (di[a] < di[best]) best = a;  // np.argmin: first minimum
        cA[i] = best;
    }
}

__global__ void dad_ijs_kernel(const int64_t *__restrict__ ij, int64_t n,
                               const double *__restrict__ D, int na,
                               const int32_t *__restrict__ cA, double *__restrict__ dad)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = ij[2 * p], j = ij[2 * p + 1];
        dad[p] = (D[i * na + cA[j]] + D[j * na + cA[i]]) / 2;
    }
}

// --- update_bounds / get_bounds_alt (annchor/utils.py:304-352): merge-join of two sorted
//     known-distance lists; every common third point k gives |d_ik - d_jk| <= d_ij <= d_ik + d_jk
__global__ void update_bounds_kernel(const int64_t *__restrict__ ij, int64_t n,
                                     const int64_t *__restrict__ kptr,
                                     const int64_t *__restrict__ kids,
                                     const double *__restrict__ kds, double *__restrict__ bounds)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = ij[2 * t], j = ij[2 * t + 1];
        int64_t p = kptr[i], pe = kptr[i + 1], q = kptr[j], qe = kptr[j + 1];
        double ub = INFINITY, lb = 0.0;
        while (p < pe && q < qe) {
            const int64_t a = kids[p], b = kids[q];
            if (a < b) ++p;
            else if (a > b) ++q;
            else {
                const double x = kds[p], y = kds[q];
                ub = fmin(ub, x + y);
                lb = fmax(lb, fabs(x - y));
                ++p;
                ++q;
            }
        }
        bounds[2 * t] = lb;
        bounds[2 * t + 1] = ub;
    }
}

// --- SimpleStratifiedLinearRegression.predict (annchor/regressors.py:71-103) + clip
//     (annchor/annchor.py:359-363).  Bin b owns (bins[b], bins[b+1]].
__global__ void predict_kernel(const double *__restrict__ feat, int64_t n,
                               const double *__restrict__ bins, const double *__restrict__ coef,
                               const double *__restrict__ icpt, int nb, double *__restrict__ raw,
                               double *__restrict__ clipped)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x) {
        const double lb = feat[4 * p], ub = feat[4 * p + 1], dad = feat[4 * p + 2];
        double y = 0.0;
        for (int b = 0; b < nb; ++b)
            if (dad > bins[b] && dad <= bins[b + 1])
                y = lb * coef[3 * b] + ub * coef[3 * b + 1] + dad * coef[3 * b + 2] + icpt[b];
        if (raw) raw[p] = y;
        if (clipped) clipped[p] = fmin(fmax(y, lb), ub);
    }
}

// --- SimpleStratifiedErrorRegression.predict (annchor/error_predictors.py:56-67): closed
//     intervals, later bins overwrite earlier ones on shared edges.
__global__ void error_labels_kernel(const double *__restrict__ f, int64_t n,
                                    const double *__restrict__ bins, int nb,
                                    int64_t *__restrict__ labels)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x) {
        const double v = f[p];
        int64_t lab = -1;
        for (int b = 0; b < nb; ++b)
            if (v >= bins[b] && v <= bins[b + 1]) lab = b;
        labels[p] = lab;
    }
}

// --- get_probs (annchor/utils.py:581-589): searchsorted(errs[label], p, 'left') / len ---
__global__ void probs_kernel(const double *__restrict__ pv, const int64_t *__restrict__ labels,
                             int64_t n, const double *__restrict__ errs,
                             const int64_t *__restrict__ eptr, double *__restrict__ prob)
{
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t l = labels[t];
        const double *e = errs + eptr[l];
        const int64_t len = eptr[l + 1] - eptr[l];
        const double x = pv[t];
        int64_t lo = 0, hi = len;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (e[mid] < x) lo = mid + 1;
            else hi = mid;
        }
        prob[t] = (double)lo / (double)len;
    }
}

// order-preserving map double -> uint64
__device__ __forceinline__ uint64_t f64_key(double v)
{
    uint64_t u = (uint64_t)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(uint64_t k)
{
    uint64_t u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)u);
}

// --- thresh loop (annchor/annchor.py:399-404): k-th smallest per CSR row, exact, by 8-bit
//     MSD radix select; one block per row.
__global__ void __launch_bounds__(256)
row_kth_kernel(const double *__restrict__ RA, const int64_t *__restrict__ row_ptr,
               const int64_t *__restrict__ row_pairs, int64_t nx, int64_t k,
               double *__restrict__ out)
{
    __shared__ unsigned hist[256];
    __shared__ uint64_t s_prefix;
    __shared__ int64_t s_k;
    for (int64_t row = blockIdx.x; row < nx; row += gridDim.x) {
        const int64_t beg = row_ptr[row], m = row_ptr[row + 1] - beg;
        if (m == 0) {
            if (threadIdx.x == 0) out[row] = NAN;
            continue;
        }
        if (threadIdx.x == 0) {
            s_prefix = 0;
            s_k = k < m ? k : m - 1;
        }
        for (int shift = 56; shift >= 0; shift -= 8) {
            hist[threadIdx.x] = 0;
            __syncthreads();
            const uint64_t prefix = s_prefix;
            const uint64_t mask = shift == 56 ? 0ull : (~0ull << (shift + 8));
            for (int64_t t = threadIdx.x; t < m; t += blockDim.x) {
                const uint64_t key = f64_key(RA[row_pairs[beg + t]]);
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xff], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                int64_t kk = s_k;
                int b = 0;
                for (; b < 256; ++b) {
                    if (kk < (int64_t)hist[b]) break;
                    kk -= hist[b];
                }
                s_k = kk;
                s_prefix = prefix | ((uint64_t)b << shift);
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) out[row] = key_f64(s_prefix);
        __syncthreads();
    }
}

// --- get_nn (annchor/utils.py:383-429): per row push not-computed entries out by +max(d),
//     emit the nn-1 smallest by (value, pair index).  One block per row, nn-1 extract-min
//     sweeps (rows are short-lived parity-mode objects; the streaming engine has its own top-k).
struct MinKey {
    double d;
    int64_t pair;
};
__device__ __forceinline__ bool key_less(const MinKey &a, const MinKey &b)
{
    return a.d < b.d || (a.d == b.d && a.pair < b.pair);
}

__global__ void __launch_bounds__(256)
get_nn_kernel(int64_t nx, int nn, const double *__restrict__ RA, const int64_t *__restrict__ ij,
              const int64_t *__restrict__ row_ptr, const int64_t *__restrict__ row_pairs,
              const uint8_t *__restrict__ ncm, int64_t *__restrict__ ngi, double *__restrict__ ngd)
{
    __shared__ double s_red[8];
    __shared__ MinKey s_key[8];
    __shared__ MinKey s_prev;
    __shared__ double s_mx;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t row = blockIdx.x; row < nx; row += gridDim.x) {
        const int64_t beg = row_ptr[row], m = row_ptr[row + 1] - beg;
        double mx = -INFINITY;
        for (int64_t t = threadIdx.x; t < m; t += blockDim.x) mx = fmax(mx, RA[row_pairs[beg + t]]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if (lane == 0) s_red[w] = mx;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int q = 1; q < 8; ++q) mx = fmax(mx, s_red[q]);
            s_mx = mx;
            s_prev = MinKey{-INFINITY, -1};
        }
        __syncthreads();
        mx = s_mx;
        for (int r = 0; r < nn - 1; ++r) {
            const MinKey prev = s_prev;
            MinKey best{INFINITY, INT64_MAX};
            for (int64_t t = threadIdx.x; t < m; t += blockDim.x) {
                const int64_t pr = row_pairs[beg + t];
                MinKey cur{RA[pr] + (ncm[pr] ? mx : 0.0), pr};
                if (key_less(prev, cur) && key_less(cur, best)) best = cur;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                MinKey y;
                y.d = __shfl_xor_sync(0xffffffffu, best.d, o);
                y.pair = __shfl_xor_sync(0xffffffffu, best.pair, o);
                if (key_less(y, best)) best = y;
            }
            if (lane == 0) s_key[w] = best;
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int q = 1; q < 8; ++q)
                    if (key_less(s_key[q], best)) best = s_key[q];
                s_prev = best;
                const int64_t o = row * (nn - 1) + r;
                if (best.pair == INT64_MAX) {  // row shorter than nn-1
                    ngd[o] = NAN;
                    ngi[o] = -1;
                } else {
                    ngd[o] = RA[best.pair];
                    ngi[o] = ij[2 * best.pair] == row ? ij[2 * best.pair + 1] : ij[2 * best.pair];
                }
            }
            __syncthreads();
        }
    }
}

static int up(annb_ctx *c, DevBuf &b, const void *src, size_t bytes)
{
    ANNB_TRY(b.ensure(bytes ? bytes : 8));
    if (bytes) ANNB_CUDA(cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return ANNB_OK;
}
static int down(annb_ctx *c, void *dst, const DevBuf &b, size_t bytes)
{
    if (bytes) ANNB_CUDA(cudaMemcpyAsync(dst, b.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    return ANNB_OK;
}
static int grid1d(const annb_ctx *c, int64_t n, int per_block)
{
    int64_t g = (n + per_block - 1) / per_block;
    int64_t cap = (int64_t)c->num_sms * 8;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

}  // namespace annb

using namespace annb;

#define ANNB_ENTER(c)                                           \
    ANNB_REQUIRE((c) != nullptr, ANNB_EINVAL, "ctx is NULL");   \
    ANNB_CUDA(cudaSetDevice((c)->device))

ANNB_API int annb_bounds_ijs(annb_ctx *c, const int64_t *ij, int64_t n, const double *D,
                             int64_t nx, int64_t na, double *bounds)
{
    ANNB_ENTER(c);
    ANNB_REQUIRE(n >= 0 && nx > 0 && na > 0, ANNB_EINVAL, "bad sizes");
    if (n == 0) return ANNB_OK;
    ANNB_REQUIRE(ij && D && bounds, ANNB_EINVAL, "NULL buffer");
    ANNB_TRY(up(c, c->s_in[0], ij, (size_t)n * 16));
    ANNB_TRY(up(c, c->s_in[1], D, (size_t)nx * na * 8));
    ANNB_TRY(c->s_out[0].ensure((size_t)n * 16));
    ANNB_LAUNCH(bounds_ijs_kernel, grid1d(c, n, 8), 256, 0, c->stream, c->s_in[0].as<int64_t>(), n,
                c->s_in[1].as<double>(), (int)na, c->s_out[0].as<double>());
    return down(c, bounds, c->s_out[0], (size_t)n * 16);
}

ANNB_API int annb_dad_ijs(annb_ctx *c, const int64_t *ij, int64_t n, const double *D, int64_t nx,
                          int64_t na, double *dad)
{
    ANNB_ENTER(c);
    ANNB_REQUIRE(n >= 0 && nx > 0 && na > 0, ANNB_EINVAL, "bad sizes");
    if (n == 0) return ANNB_OK;
    ANNB_REQUIRE(ij && D && dad, ANNB_EINVAL, "NULL buffer");
    ANNB_TRY(up(c, c->s_in[0], ij, (size_t)n * 16));
    ANNB_TRY(up(c, c->s_in[1], D, (size_t)nx * na * 8));
    ANNB_TRY(c->s_in[2].ensure((size_t)nx * 4));
    ANNB_TRY(c->s_out[0].ensure((size_t)n * 8));
    ANNB_LAUNCH(closest_anchor_kernel, grid1d(c, nx, 256), 256, 0, c->stream,
                c->s_in[1].as<double>(), nx, (int)na, c->s_in[2].as<int32_t>());
    ANNB_LAUNCH(dad_ijs_kernel, grid1d(c, n, 256), 256, 0, c->stream, c->s_in[0].as<int64_t>(), n,
                c->s_in[1].as<double>(), (int)na, c->s_in[2].as<int32_t>(),
                c->s_out[0].as<double>());
    return down(c, dad, c->s_out[0], (size_t)n * 8);
}

ANNB_API int annb_update_bounds(annb_ctx *c, const int64_t *ij, int64_t n, const int64_t *kptr,
                                const int64_t *kids, const double *kds, int64_t nx, double *bounds)
{
    ANNB_ENTER(c);
    ANNB_REQUIRE(n >= 0 && nx > 0, ANNB_EINVAL, "bad sizes");
    if (n == 0) return ANNB_OK;
    ANNB_REQUIRE(ij && kptr && bounds, ANNB_EINVAL, "NULL buffer");
    const int64_t nk = kptr[nx];
    ANNB_TRY(up(c, c->s_in[0], ij, (size_t)n * 16));
    ANNB_TRY(up(c, c->s_in[1], kptr, (size_t)(nx + 1) * 8));
    ANNB_TRY(up(c, c->s_in[2], kids, (size_t)nk * 8));
    ANNB_TRY(up(c, c->s_in[3], kds, (size_t)nk * 8));
    ANNB_TRY(c->s_out[0].ensure((size_t)n * 16));
    ANNB_LAUNCH(update_bounds_kernel, grid1d(c, n, 128), 128, 0, c->stream,
                c->s_in[0].as<int64_t>(), n, c->s_in[1].as<int64_t>(), c->s_in[2].as<int64_t>(),
                c->s_in[3].as<double>(), c->s_out[0].as<double>());
    return down(c, bounds, c->s_out[0], (size_t)n * 16);
}

ANNB_API int annb_predict_stratified(annb_ctx *c, const double *features, int64_t n,
                                     const double *bins, const double *coef, const double *icpt,
                                     int64_t nb, double *pred_raw, double *pred_clipped)
{
    ANNB_ENTER(c);
    ANNB_REQUIRE(n >= 0 && nb > 0, ANNB_EINVAL, "bad sizes");
    if (n == 0) return ANNB_OK;
    ANNB_REQUIRE(features && bins && coef && icpt, ANNB_EINVAL, "NULL buffer");
    ANNB_TRY(up(c, c->s_in[0], features, (size_t)n * 32));
    ANNB_TRY(up(c, c->s_in[1], bins, (size_t)(nb + 1) * 8));
    ANNB_TRY(up(c, c->s_in[2], coef, (size_t)nb * 24));
    ANNB_TRY(up(c, c->s_in[3], icpt, (size_t)nb * 8));
    ANNB_TRY(c->s_out[0].ensure((size_t)n * 8));
    ANNB_TRY(c->s_out[1].ensure((size_t)n * 8));
    ANNB_LAUNCH(predict_kernel, grid1d(c, n, 256), 256, 0, c->stream, c->s_in[0].as<double>(), n,
                c->s_in[1].as<double>(), c->s_in[2].as<double>(), c->s_in[3].as<double>(), (int)nb,
                c->s_out[0].as<double>(), c->s_out[1].as<double>());
    if (pred_raw)
        ANNB_CUDA(cudaMemcpyAsync(pred_raw, c->s_out[0].p, (size_t)n * 8, cudaMemcpyDeviceToHost,
                                  c->stream));
    return down(c, pred_clipped, c->s_out[1], pred_clipped ? (size_t)n * 8 : 0);
}

ANNB_API int annb_error_labels(annb_ctx *c, const double *feature, int64_t n, const double *bins,
                               int64_t nb, int64_t *labels)
{
    ANNB_ENTER(c);
    ANNB_REQUIRE(n >= 0 && nb > 0, ANNB_EINVAL, "bad sizes");
    if (n == 0) return ANNB_OK;
    ANNB_REQUIRE(feature && bins && labels, ANNB_EINVAL, "NULL buffer");
    ANNB_TRY(up(c, c->s_in[0], feature, (size_t)n * 8));
    ANNB_TRY(up(c, c->s_in[1], bins, (size_t)(nb + 1) * 8));
    ANNB_TRY(c->s_out[0].ensure((size_t)n * 8));
    ANNB_LAUNCH(error_labels_kernel, grid1d(c, n, 256), 256, 0, c->stream, c->s_in[0].as<double>(),
                n, c->s_in[1].as<double>(), (int)nb, c->s_out[0].as<int64_t>());
    return down(c, labels, c->s_out[0], (size_t)n * 8);
}

ANNB_API int annb_probs(annb_ctx *c, const double *p, const int64_t *labels, int64_t n,
                        const double *errs, const int64_t *eptr, int64_t nl, double *prob)
{
    ANNB_ENTER(c);
    ANNB_REQUIRE(n >= 0 && nl > 0, ANNB_EINVAL, "bad sizes");
    if (n == 0) return ANNB_OK;
    ANNB_REQUIRE(p && labels && errs && eptr && prob, ANNB_EINVAL, "NULL buffer");
    for (int64_t t = 0; t < n; ++t)
        ANNB_REQUIRE(labels[t] >= 0 && labels[t] < nl, ANNB_EINVAL, "label %lld out of range at %lld",
                     (long long)labels[t], (long long)t);
    ANNB_TRY(up(c, c->s_in[0], p, (size_t)n * 8));
    ANNB_TRY(up(c, c->s_in[1], labels, (size_t)n * 8));
    ANNB_TRY(up(c, c->s_in[2], errs, (size_t)eptr[nl] * 8));
    ANNB_TRY(up(c, c->s_in[3], eptr, (size_t)(nl + 1) * 8));
    ANNB_TRY(c->s_out[0].ensure((size_t)n * 8));
    ANNB_LAUNCH(probs_kernel, grid1d(c, n, 256), 256, 0, c->stream, c->s_in[0].as<double>(),
                c->s_in[1].as<int64_t>(), n, c->s_in[2].as<double>(), c->s_in[3].as<int64_t>(),
                c->s_out[0].as<double>());
    return down(c, prob, c->s_out[0], (size_t)n * 8);
}

ANNB_API int annb_row_kth(annb_ctx *c, const double *RA, int64_t npairs, const int64_t *row_ptr,
                          const int64_t *row_pairs, int64_t nx, int64_t k, double *out)
{
    ANNB_ENTER(c);
    ANNB_REQUIRE(nx > 0 && k >= 0 && npairs >= 0, ANNB_EINVAL, "bad sizes");
    ANNB_REQUIRE(RA && row_ptr && row_pairs && out, ANNB_EINVAL, "NULL buffer");
    ANNB_TRY(up(c, c->s_in[0], RA, (size_t)npairs * 8));
    ANNB_TRY(up(c, c->s_in[1], row_ptr, (size_t)(nx + 1) * 8));
    ANNB_TRY(up(c, c->s_in[2], row_pairs, (size_t)row_ptr[nx] * 8));
    ANNB_TRY(c->s_out[0].ensure((size_t)nx * 8));
    int grid = (int)(nx < (int64_t)c->num_sms * 8 ? nx : (int64_t)c->num_sms * 8);
    ANNB_LAUNCH(row_kth_kernel, grid, 256, 0, c->stream, c->s_in[0].as<double>(),
                c->s_in[1].as<int64_t>(), c->s_in[2].as<int64_t>(), nx, k, c->s_out[0].as<double>());
    return down(c, out, c->s_out[0], (size_t)nx * 8);
}

ANNB_API int annb_get_nn(annb_ctx *c, int64_t nx, int64_t nn, const double *RA, const int64_t *ij,
                         int64_t npairs, const int64_t *row_ptr, const int64_t *row_pairs,
                         const uint8_t *not_computed, int64_t *ngi, double *ngd)
{
    ANNB_ENTER(c);
    ANNB_REQUIRE(nx > 0 && nn >= 2 && npairs >= 0, ANNB_EINVAL, "bad sizes");
    ANNB_REQUIRE(RA && ij && row_ptr && row_pairs && not_computed && ngi && ngd, ANNB_EINVAL,
                 "NULL buffer");
    ANNB_TRY(up(c, c->s_in[0], RA, (size_t)npairs * 8));
    ANNB_TRY(up(c, c->s_in[1], ij, (size_t)npairs * 16));
    ANNB_TRY(up(c, c->s_in[2], row_ptr, (size_t)(nx + 1) * 8));
    ANNB_TRY(up(c, c->s_in[3], row_pairs, (size_t)row_ptr[nx] * 8));
    ANNB_TRY(up(c, c->s_in[4], not_computed, (size_t)npairs));
    ANNB_TRY(c->s_out[0].ensure((size_t)nx * (nn - 1) * 8));
    ANNB_TRY(c->s_out[1].ensure((size_t)nx * (nn - 1) * 8));
    int grid = (int)(nx < (int64_t)c->num_sms * 8 ? nx : (int64_t)c->num_sms * 8);
    ANNB_LAUNCH(get_nn_kernel, grid, 256, 0, c->stream, nx, (int)nn, c->s_in[0].as<double>(),
                c->s_in[1].as<int64_t>(), c->s_in[2].as<int64_t>(), c->s_in[3].as<int64_t>(),
                c->s_in[4].as<uint8_t>(), c->s_out[0].as<int64_t>(), c->s_out[1].as<double>());
    ANNB_CUDA(cudaMemcpyAsync(ngi, c->s_out[0].p, (size_t)nx * (nn - 1) * 8, cudaMemcpyDeviceToHost,
                              c->stream));
    return down(c, ngd, c->s_out[1], (size_t)nx * (nn - 1) * 8);
}
