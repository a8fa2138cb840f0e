// metrics.cu -- K4 (batched pair distances == get_exact_ijs, annchor/utils.py:110-177) and
// the one-anchor-to-all rows K1 is built from (annchor/pickers.py:45-46), for the bundled
// metrics: euclidean (annchor/distances.py:8-13), cosine (annchor/utils.py:14,67),
// levenshtein (annchor/distances.py:16-20), 1-D wasserstein (annchor/utils.py:75-86).
//
// Rooflines: dense gathers are HBM/L2 bound (2 rows of d elements per pair, read with
// 128-bit coalesced loads, one warp per pair); Levenshtein is integer-ALU bound
// (Myers/Hyyro bit-parallel DP, one thread per pair, pattern bit-tables in shared memory).
#include "common.cuh"

namespace annb {

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------
// dense rows: euclidean / cosine
// ---------------------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float> {
    using type = float4;
    static constexpr int N = 4;
};
template <> struct Vec<double> {
    using type = double2;
    static constexpr int N = 2;
};

template <typename T>
__device__ __forceinline__ void load_vec(const T *p, T (&v)[Vec<T>::N])
{
    typename Vec<T>::type q = __ldg(reinterpret_cast<const typename Vec<T>::type *>(p));
    const T *s = reinterpret_cast<const T *>(&q);
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) v[k] = s[k];
}

template <typename T, int METRIC>
__device__ __forceinline__ void dense_accum(const T (&a)[Vec<T>::N], const T (&b)[Vec<T>::N],
                                            double &s0, double &s1, double &s2)
{
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) {
        if (METRIC == ANNB_EUCLIDEAN) {
            T t = a[k] - b[k];  // difference in the input precision, like x - y in the reference
            s0 += (double)t * (double)t;
        } else {
            s0 += (double)a[k] * (double)b[k];
            s1 += (double)a[k] * (double)a[k];
            s2 += (double)b[k] * (double)b[k];
        }
    }
}

template <typename T, int METRIC>
__device__ __forceinline__ double dense_finish(double s0, double s1, double s2)
{
    s0 = warp_sum(s0);
    if (METRIC == ANNB_EUCLIDEAN) {
        double r = sqrt(s0);
        // np.linalg.norm on float32 input returns float32 precision (annchor/utils.py:146-149)
        return sizeof(T) == 4 ? (double)(float)r : r;
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    double r = 1.0 - s0 / sqrt(s1 * s2);
    return fmin(fmax(r, 0.0), 2.0);
}

// one warp per pair, UNROLL pairs in flight per warp so 2*UNROLL 128-bit loads are
// outstanding per lane before the first use.
template <typename T, int METRIC, typename OutT, int UNROLL>
__global__ void __launch_bounds__(256)
dense_pair_kernel(const T *__restrict__ X, int64_t ld, const int32_t *__restrict__ I,
                  const int32_t *__restrict__ J, int64_t n, OutT *__restrict__ out)
{
    constexpr int VN = Vec<T>::N;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * UNROLL; base < n; base += nwarps * UNROLL) {
        double s0[UNROLL], s1[UNROLL], s2[UNROLL];
        const T *xi[UNROLL], *xj[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            int64_t p = base + u < n ? base + u : n - 1;
            xi[u] = X + (int64_t)__ldg(I + p) * ld;
            xj[u] = X + (int64_t)__ldg(J + p) * ld;
            s0[u] = s1[u] = s2[u] = 0.0;
        }
        for (int64_t k = lane * VN; k < ld; k += 32 * VN) {
            T a[UNROLL][VN], b[UNROLL][VN];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                load_vec<T>(xi[u] + k, a[u]);
                load_vec<T>(xj[u] + k, b[u]);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                dense_accum<T, METRIC>(a[u], b[u], s0[u], s1[u], s2[u]);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            double r = dense_finish<T, METRIC>(s0[u], s1[u], s2[u]);
            if (lane == 0 && base + u < n) out[base + u] = (OutT)r;
        }
    }
}

// distances from item *anchor to every item: anchor row staged in shared memory once per
// block, X streamed with coalesced 128-bit loads (N*d*sizeof(T) bytes per launch).
template <typename T, int METRIC>
__global__ void __launch_bounds__(256)
dense_anchor_kernel(const T *__restrict__ X, int64_t ld, int64_t n,
                    const int32_t *__restrict__ anchor, double *__restrict__ row)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *sa = reinterpret_cast<T *>(smem_raw);
    constexpr int VN = Vec<T>::N;
    const T *xa = X + (int64_t)(*anchor) * ld;
    for (int64_t k = threadIdx.x; k < ld; k += blockDim.x) sa[k] = xa[k];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * 2; base < n; base += nwarps * 2) {
        double s0[2] = {0, 0}, s1[2] = {0, 0}, s2[2] = {0, 0};
        const T *xj[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) xj[u] = X + (base + u < n ? base + u : n - 1) * ld;
        for (int64_t k = lane * VN; k < ld; k += 32 * VN) {
            T a[VN], b[2][VN];
#pragma unroll
            for (int u = 0; u < 2; ++u) load_vec<T>(xj[u] + k, b[u]);
#pragma unroll
            for (int q = 0; q < VN; ++q) a[q] = sa[k + q];
#pragma unroll
            for (int u = 0; u < 2; ++u) dense_accum<T, METRIC>(a, b[u], s0[u], s1[u], s2[u]);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            double r = dense_finish<T, METRIC>(s0[u], s1[u], s2[u]);
            if (lane == 0 && base + u < n) row[base + u] = r;
        }
    }
}

// ---------------------------------------------------------------------------------
// 1-D Wasserstein over precomputed unit-mass CDFs: sum_b |CDF_i(b) - CDF_j(b)|
// ---------------------------------------------------------------------------------
template <typename OutT>
__global__ void __launch_bounds__(256)
w1_pair_kernel(const double *__restrict__ C, int64_t nb, const int32_t *__restrict__ I,
               const int32_t *__restrict__ J, const int32_t *__restrict__ anchor, int64_t n,
               OutT *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = warp; p < n; p += nwarps) {
        const double *ci = C + (int64_t)(anchor ? *anchor : __ldg(I + p)) * nb;
        const double *cj = C + (int64_t)(anchor ? (int32_t)p : __ldg(J + p)) * nb;
        double s = 0.0;
        for (int64_t k = lane; k < nb; k += 32) s += fabs(__ldg(ci + k) - __ldg(cj + k));
        s = warp_sum(s);
        if (lane == 0) out[p] = (OutT)s;
    }
}

// ---------------------------------------------------------------------------------
// General cost-matrix Wasserstein (annchor/utils.py:75-86: kantorovich(x, y, cost=M) of pynndescent
// 0.5.13 -- zero bins dropped, both histograms normalised to unit mass, exact optimal transport).
// One warp per pair solves the transportation problem exactly by successive shortest augmenting
// paths with node potentials (multi-source Dijkstra on the dense residual graph, reduced costs >= 0):
// every augmentation exhausts a supply, a demand or a flow arc.  Supports up to 64 bins (the
// reference's 8 x 8 digits); flows, potentials and distances are float64 in shared memory.
// ---------------------------------------------------------------------------------
constexpr int OT_B = 64;
struct OtWarp {
    double f[OT_B * OT_B];  // flow on (source k, sink l), row pitch = n
    double ds[OT_B], dt[OT_B], ps[OT_B], pt[OT_B], ra[OT_B], rb[OT_B];
    int16_t si[OT_B], tj[OT_B], prev_s[OT_B], prev_t[OT_B];
    uint8_t vis_s[OT_B], vis_t[OT_B];
};

__device__ __forceinline__ double shfl_xor_f64(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

// mass: unit-sum histograms (n_items, nb); cost: (nb, nb) row-major; returns the optimal transport cost
__device__ double ot_solve(OtWarp &W, const double *__restrict__ sC /* shared, nb x nb */, int nb,
                           const double *__restrict__ x, const double *__restrict__ y, int lane)
{
    constexpr double EPS = 1e-13, TINY = 1e-15;
    // 1. compress the supports
    int m = 0, n = 0;
    for (int b0 = 0; b0 < nb; b0 += 32) {
        const int b = b0 + lane;
        const double xv = b < nb ? x[b] : 0.0, yv = b < nb ? y[b] : 0.0;
        const unsigned mx = __ballot_sync(0xffffffffu, xv > 0.0), my = __ballot_sync(0xffffffffu, yv > 0.0);
        if (xv > 0.0) {
            const int k = m + __popc(mx & ((1u << lane) - 1));
            W.si[k] = (int16_t)b;
            W.ra[k] = xv;
        }
        if (yv > 0.0) {
            const int l = n + __popc(my & ((1u << lane) - 1));
            W.tj[l] = (int16_t)b;
            W.rb[l] = yv;
        }
        m += __popc(mx);
        n += __popc(my);
    }
    __syncwarp();
    if (m == 0 || n == 0) return 0.0;
    for (int q = lane; q < m * n; q += 32) W.f[q] = 0.0;
    for (int k = lane; k < m; k += 32) W.ps[k] = 0.0;
    for (int l = lane; l < n; l += 32) {
        double mn = INFINITY;
        for (int k = 0; k < m; ++k) mn = fmin(mn, sC[W.si[k] * nb + W.tj[l]]);
        W.pt[l] = mn;  // reduced cost C + ps - pt >= 0 on every forward arc
    }
    __syncwarp();
    // 2. augment until all supply is shipped
    for (int round = 0; round < 8 * (m + n) + 32; ++round) {
        double rem = 0.0;
        for (int k = lane; k < m; k += 32) rem += W.ra[k] > EPS ? W.ra[k] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rem += shfl_xor_f64(rem, o);
        if (!(rem > EPS)) break;
        for (int k = lane; k < m; k += 32) {
            W.ds[k] = W.ra[k] > EPS ? 0.0 : INFINITY;
            W.prev_s[k] = -1;
            W.vis_s[k] = 0;
        }
        for (int l = lane; l < n; l += 32) {
            W.dt[l] = INFINITY;
            W.prev_t[l] = -1;
            W.vis_t[l] = 0;
        }
        __syncwarp();
        int target = -1;
        double dtar = INFINITY;
        for (int it = 0; it < m + n; ++it) {
            // closest unvisited node (sources: id k, sinks: id 64 + l)
            double bv = INFINITY;
            int bid = 1 << 20;
            for (int k = lane; k < m; k += 32)
                if (!W.vis_s[k] && W.ds[k] < bv) {
                    bv = W.ds[k];
                    bid = k;
                }
            for (int l = lane; l < n; l += 32)
                if (!W.vis_t[l] && (W.dt[l] < bv)) {
                    bv = W.dt[l];
                    bid = OT_B + l;
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = shfl_xor_f64(bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bid, o);
                if (ov < bv || (ov == bv && oi < bid)) {
                    bv = ov;
                    bid = oi;
                }
            }
            if (!(bv < INFINITY)) break;
            if (bid >= OT_B) {
                const int l = bid - OT_B;
                if (lane == 0) W.vis_t[l] = 1;
                if (W.rb[l] > EPS) {
                    target = l;
                    dtar = bv;
                    break;
                }
                // backward arcs l -> k where flow can be withdrawn: cost -C, reduced -C + pt - ps
                for (int k = lane; k < m; k += 32)
                    if (!W.vis_s[k] && W.f[k * n + l] > TINY) {
                        const double rc = fmax(W.pt[l] - W.ps[k] - sC[W.si[k] * nb + W.tj[l]], 0.0);
                        if (bv + rc < W.ds[k]) {
                            W.ds[k] = bv + rc;
                            W.prev_s[k] = (int16_t)l;
                        }
                    }
            } else {
                const int k = bid;
                if (lane == 0) W.vis_s[k] = 1;
                for (int l = lane; l < n; l += 32)
                    if (!W.vis_t[l]) {
                        const double rc = fmax(sC[W.si[k] * nb + W.tj[l]] + W.ps[k] - W.pt[l], 0.0);
                        if (bv + rc < W.dt[l]) {
                            W.dt[l] = bv + rc;
                            W.prev_t[l] = (int16_t)k;
                        }
                    }
            }
            __syncwarp();
        }
        __syncwarp();
        if (target < 0) break;  // (cannot happen with positive remaining supply and demand)
        // potentials: visited nodes move by their distance, the others by the target's
        for (int k = lane; k < m; k += 32) W.ps[k] += W.vis_s[k] ? W.ds[k] : dtar;
        for (int l = lane; l < n; l += 32) W.pt[l] += W.vis_t[l] ? W.dt[l] : dtar;
        __syncwarp();
        if (lane == 0) {
            // bottleneck along target -> ... -> root (a source with remaining supply)
            double amount = W.rb[target];
            int l = target, k;
            for (;;) {
                k = W.prev_t[l];
                if (W.prev_s[k] < 0) break;
                const int l2 = W.prev_s[k];
                amount = fmin(amount, W.f[k * n + l2]);
                l = l2;
            }
            amount = fmin(amount, W.ra[k]);
            const int root = k;
            l = target;
            for (;;) {
                k = W.prev_t[l];
                W.f[k * n + l] += amount;
                if (W.prev_s[k] < 0) break;
                const int l2 = W.prev_s[k];
                const double r = W.f[k * n + l2] - amount;
                W.f[k * n + l2] = r > TINY ? r : 0.0;
                l = l2;
            }
            W.ra[root] -= amount;
            W.rb[target] -= amount;
        }
        __syncwarp();
    }
    double cost = 0.0;
    for (int q = lane; q < m * n; q += 32) cost += W.f[q] * sC[W.si[q / n] * nb + W.tj[q % n]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cost += shfl_xor_f64(cost, o);
    return cost;
}

constexpr int OT_WARPS = 4;
template <typename OutT>
__global__ void __launch_bounds__(OT_WARPS * 32)
ot_pair_kernel(const double *__restrict__ mass, int nb, const double *__restrict__ cost,
               const int32_t *__restrict__ I, const int32_t *__restrict__ J, const int32_t *__restrict__ anchor,
               const int32_t *__restrict__ perm, int64_t n, OutT *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char ot_smem[];
    double *sC = reinterpret_cast<double *>(ot_smem);
    OtWarp *ws = reinterpret_cast<OtWarp *>(sC + OT_B * OT_B);
    for (int q = threadIdx.x; q < nb * nb; q += blockDim.x) sC[q] = cost[q];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    OtWarp &W = ws[wib];
    const int64_t warp = (int64_t)blockIdx.x * OT_WARPS + wib, nwarps = (int64_t)gridDim.x * OT_WARPS;
    for (int64_t q = warp; q < n; q += nwarps) {
        const int64_t p = perm ? (int64_t)__ldg(perm + q) : q;
        const int64_t a = anchor ? *anchor : __ldg(I + p), b = anchor ? p : __ldg(J + p);
        const double r = a == b ? 0.0 : ot_solve(W, sC, nb, mass + a * nb, mass + b * nb, lane);
        if (lane == 0) out[p] = (OutT)r;
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Levenshtein: Myers (1999) / Hyyro (2003) bit-parallel unit-cost edit distance, global
// alignment, multi-word with +-1 horizontal carries between 64-row blocks.  One thread
// per pair; the pattern's match bit-tables Peq[symbol][word] are built once per warp in
// shared memory and shared by every lane whose pair has that pattern (pairs arrive
// grouped by pattern: anchor rows trivially, refine lists via group_pairs_by_i()).
// ---------------------------------------------------------------------------------
template <int W>
__device__ __forceinline__ int lev_thread(const uint64_t *peq, int weff, int m,
                                          const uint8_t *__restrict__ text, int n)
{
    if (m == 0) return n;
    uint64_t Pv[W], Mv[W];
#pragma unroll
    for (int w = 0; w < W; ++w) {
        Pv[w] = ~0ull;
        Mv[w] = 0ull;
    }
    int score = m;
    const int last_w = (m - 1) >> 6;
    const uint64_t last_bit = 1ull << ((m - 1) & 63);
    for (int t0 = 0; t0 < n; t0 += 16) {
        const uint4 chunk = __ldg(reinterpret_cast<const uint4 *>(text + t0));
        const uint32_t cw[4] = {chunk.x, chunk.y, chunk.z, chunk.w};
        const int tn = min(16, n - t0);
        for (int t = 0; t < tn; ++t) {
            const int c = (cw[t >> 2] >> ((t & 3) * 8)) & 0xff;
            const uint64_t *pc = peq + c * weff;
            int hin = 1;  // D[0][j] - D[0][j-1] = +1 (global alignment)
#pragma unroll
            for (int w = 0; w < W; ++w) {
                if (w < weff) {
                    uint64_t eq = pc[w];
                    const uint64_t pv = Pv[w], mv = Mv[w];
                    const uint64_t xv = eq | mv;
                    eq |= (uint64_t)(hin < 0);
                    const uint64_t xh = (((eq & pv) + pv) ^ pv) | eq;
                    uint64_t ph = mv | ~(xh | pv);
                    uint64_t mh = pv & xh;
                    int hout;
                    if (w == last_w) {
                        hout = ((ph & last_bit) != 0) - ((mh & last_bit) != 0);
                        score += hout;
                    } else {
                        hout = (int)(ph >> 63) - (int)(mh >> 63);
                    }
                    ph = (ph << 1) | (uint64_t)(hin > 0);
                    mh = (mh << 1) | (uint64_t)(hin < 0);
                    Pv[w] = mh | ~(xv | ph);
                    Mv[w] = ph & xv;
                    hin = hout;
                }
            }
        }
    }
    return score;
}

// build Peq for pattern `pi` into this warp's shared-memory table: lane w owns word w.
__device__ __forceinline__ void lev_build_peq(uint64_t *peq, int sigma, int weff,
                                              const uint8_t *__restrict__ pat, int m, int lane)
{
    for (int k = lane; k < sigma * weff; k += 32) peq[k] = 0ull;
    __syncwarp();
    for (int w = lane; w < weff; w += 32) {
        const int lo = w * 64, hi = min(m, lo + 64);
        for (int pos = lo; pos < hi; ++pos) peq[pat[pos] * weff + w] |= 1ull << (pos - lo);
    }
    __syncwarp();
}

template <int W, typename OutT>
__global__ void __launch_bounds__(128)
lev_pair_kernel(const uint8_t *__restrict__ sym, const int64_t *__restrict__ offs,
                const int32_t *__restrict__ lens, int sigma, int weff_max,
                const int32_t *__restrict__ I, const int32_t *__restrict__ J,
                const int32_t *__restrict__ anchor, const int32_t *__restrict__ perm, int64_t n,
                OutT *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    uint64_t *peq = reinterpret_cast<uint64_t *>(smem_raw) + (size_t)wib * sigma * weff_max;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * 32; base < n; base += nwarps * 32) {
        const int64_t q = base + lane;
        const bool valid = q < n;
        const int64_t p = valid ? (perm ? (int64_t)__ldg(perm + q) : q) : 0;
        int pi = 0, tj = 0;
        if (valid) {
            pi = anchor ? *anchor : __ldg(I + p);
            tj = anchor ? (int32_t)p : __ldg(J + p);
        }
        unsigned remaining = __ballot_sync(0xffffffffu, valid);
        while (remaining) {
            const int leader = __ffs(remaining) - 1;
            const int cur = __shfl_sync(0xffffffffu, pi, leader);
            const unsigned grp = __ballot_sync(0xffffffffu, valid && pi == cur) & remaining;
            const int m = __ldg(lens + cur);
            const int weff = (m + 63) >> 6;
            lev_build_peq(peq, sigma, weff, sym + __ldg(offs + cur), m, lane);
            if ((grp >> lane) & 1u) {
                const int r = lev_thread<W>(peq, weff, m, sym + __ldg(offs + tj), __ldg(lens + tj));
                out[p] = (OutT)r;
            }
            __syncwarp();
            remaining &= ~grp;
        }
    }
}

// ---------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------
static int grid_for(const annb_ctx *c, int64_t work_items, int per_block, int blocks_per_sm)
{
    int64_t need = (work_items + per_block - 1) / per_block;
    int64_t cap = (int64_t)c->num_sms * blocks_per_sm;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

template <typename T, typename OutT>
static int launch_dense_pairs(annb_ctx *c, const annb_dataset *ds, int metric, const int32_t *I,
                              const int32_t *J, int64_t n, OutT *out)
{
    const T *X = (const T *)ds->data;
    int grid = grid_for(c, n, 8 * 4, 8);
    if (metric == ANNB_EUCLIDEAN)
        ANNB_LAUNCH((dense_pair_kernel<T, ANNB_EUCLIDEAN, OutT, 4>), grid, 256, 0, c->stream, X,
                    ds->ld, I, J, n, out);
    else
        ANNB_LAUNCH((dense_pair_kernel<T, ANNB_COSINE, OutT, 4>), grid, 256, 0, c->stream, X,
                    ds->ld, I, J, n, out);
    return ANNB_OK;
}

template <typename OutT>
static int launch_lev(annb_ctx *c, const annb_dataset *ds, const int32_t *I, const int32_t *J,
                      const int32_t *anchor, const int32_t *perm, int64_t n, OutT *out)
{
    const int weff_max = (int)((ds->max_len + 63) / 64) > 0 ? (int)((ds->max_len + 63) / 64) : 1;
    ANNB_REQUIRE(weff_max <= 16, ANNB_ERANGE,
                 "levenshtein: longest string has %lld symbols; this build supports <= 1024",
                 (long long)ds->max_len);
    const int warps = 4;
    size_t smem = (size_t)warps * ds->sigma * weff_max * sizeof(uint64_t);
    ANNB_REQUIRE(smem <= 200 * 1024, ANNB_ERANGE,
                 "levenshtein: alphabet %d x %d words does not fit shared memory", ds->sigma,
                 weff_max);
    int grid = grid_for(c, n, warps * 32, 8);
    const uint8_t *sym = (const uint8_t *)ds->data;
#define ANNB_LEV_CASE(WW)                                                                       \
    do {                                                                                        \
        if (smem > 48 * 1024)                                                                   \
            ANNB_CUDA(cudaFuncSetAttribute(lev_pair_kernel<WW, OutT>,                           \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                           (int)smem));                                         \
        ANNB_LAUNCH((lev_pair_kernel<WW, OutT>), grid, warps * 32, smem, c->stream, sym,        \
                    ds->offs, ds->lens, ds->sigma, weff_max, I, J, anchor, perm, n, out);       \
    } while (0)
    if (weff_max <= 1) ANNB_LEV_CASE(1);
    else if (weff_max <= 2) ANNB_LEV_CASE(2);
    else if (weff_max <= 4) ANNB_LEV_CASE(4);
    else if (weff_max <= 8) ANNB_LEV_CASE(8);
    else if (weff_max <= 10) ANNB_LEV_CASE(10);
    else ANNB_LEV_CASE(16);
#undef ANNB_LEV_CASE
    return ANNB_OK;
}

template <typename OutT>
static int launch_ot(annb_ctx *c, const annb_dataset *ds, const int32_t *I, const int32_t *J, const int32_t *anchor,
                     const int32_t *perm, int64_t n, OutT *out)
{
    const size_t smem = (size_t)OT_B * OT_B * 8 + (size_t)OT_WARPS * sizeof(OtWarp);
    ANNB_CUDA(cudaFuncSetAttribute(ot_pair_kernel<OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = grid_for(c, n, OT_WARPS, 1);
    ANNB_LAUNCH(ot_pair_kernel<OutT>, grid, OT_WARPS * 32, smem, c->stream, (const double *)ds->data, (int)ds->d,
                ds->cost, I, J, anchor, perm, n, out);
    return ANNB_OK;
}

template <typename OutT>
static int pair_dists_any(annb_ctx *c, const annb_dataset *ds, int metric, const int32_t *I,
                          const int32_t *J, const int32_t *perm, int64_t n, OutT *out)
{
    ANNB_TRY(check_metric(ds, metric));
    if (n == 0) return ANNB_OK;
    switch (metric) {
    case ANNB_EUCLIDEAN:
    case ANNB_COSINE:
        return ds->dtype == ANNB_F32 ? launch_dense_pairs<float, OutT>(c, ds, metric, I, J, n, out)
                                     : launch_dense_pairs<double, OutT>(c, ds, metric, I, J, n, out);
    case ANNB_LEVENSHTEIN:
        return launch_lev<OutT>(c, ds, I, J, nullptr, perm, n, out);
    case ANNB_WASSERSTEIN1D: {
        int grid = grid_for(c, n, 8, 8);
        ANNB_LAUNCH(w1_pair_kernel<OutT>, grid, 256, 0, c->stream, (const double *)ds->data, ds->d,
                    I, J, (const int32_t *)nullptr, n, out);
        return ANNB_OK;
    }
    case ANNB_WASSERSTEIN:
        return launch_ot<OutT>(c, ds, I, J, nullptr, perm, n, out);
    }
    return ANNB_EINVAL;
}

int pair_dists_f64(annb_ctx *c, const annb_dataset *ds, int metric, const int32_t *I,
                   const int32_t *J, int64_t n, double *out)
{
    return pair_dists_any<double>(c, ds, metric, I, J, nullptr, n, out);
}

int pair_dists_f32_perm(annb_ctx *c, const annb_dataset *ds, int metric, const int32_t *I,
                        const int32_t *J, const int32_t *perm, int64_t n, float *out)
{
    return pair_dists_any<float>(c, ds, metric, I, J, perm, n, out);
}

int anchor_row_f64(annb_ctx *c, const annb_dataset *ds, int metric, const int32_t *anchor,
                   double *row)
{
    ANNB_TRY(check_metric(ds, metric));
    const int64_t n = ds->n;
    switch (metric) {
    case ANNB_EUCLIDEAN:
    case ANNB_COSINE: {
        int grid = grid_for(c, n, 8 * 2, 8);
        size_t smem = (size_t)ds->ld * (ds->dtype == ANNB_F32 ? 4 : 8);
        ANNB_REQUIRE(smem <= 200 * 1024, ANNB_ERANGE, "row of %lld elements exceeds shared memory",
                     (long long)ds->d);
#define ANNB_ANCHOR_CASE(T, M)                                                                  \
    do {                                                                                        \
        if (smem > 48 * 1024)                                                                   \
            ANNB_CUDA(cudaFuncSetAttribute(dense_anchor_kernel<T, M>,                           \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                           (int)smem));                                         \
        ANNB_LAUNCH((dense_anchor_kernel<T, M>), grid, 256, smem, c->stream,                    \
                    (const T *)ds->data, ds->ld, n, anchor, row);                               \
    } while (0)
        if (ds->dtype == ANNB_F32) {
            if (metric == ANNB_EUCLIDEAN) ANNB_ANCHOR_CASE(float, ANNB_EUCLIDEAN);
            else ANNB_ANCHOR_CASE(float, ANNB_COSINE);
        } else {
            if (metric == ANNB_EUCLIDEAN) ANNB_ANCHOR_CASE(double, ANNB_EUCLIDEAN);
            else ANNB_ANCHOR_CASE(double, ANNB_COSINE);
        }
#undef ANNB_ANCHOR_CASE
        return ANNB_OK;
    }
    case ANNB_LEVENSHTEIN:
        return launch_lev<double>(c, ds, nullptr, nullptr, anchor, nullptr, n, row);
    case ANNB_WASSERSTEIN1D: {
        int grid = grid_for(c, n, 8, 8);
        ANNB_LAUNCH(w1_pair_kernel<double>, grid, 256, 0, c->stream, (const double *)ds->data,
                    ds->d, (const int32_t *)nullptr, (const int32_t *)nullptr, anchor, n, row);
        return ANNB_OK;
    }
    case ANNB_WASSERSTEIN:
        return launch_ot<double>(c, ds, nullptr, nullptr, anchor, nullptr, n, row);
    }
    return ANNB_EINVAL;
}

}  // namespace annb
