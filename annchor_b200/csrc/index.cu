// index.cu -- the streaming index: Annchor.fit() (annchor/annchor.py:532-623) without any
// Theta(N^2) array.  Host-side orchestration of the sweeps in sweep_*.cu plus the small kernels
// around them (point metadata, candidate counting, known-pair hash map, selection, top-k).
#include <algorithm>
#include <cmath>
#include <unordered_map>
#include <unordered_set>

#include "sweep.cuh"
#include "sweep_args.cuh"
#include "index_obj.cuh"

namespace annb {

// ---- declarations from the other translation units -------------------------------------------
int maxmin_device(annb_ctx *c, const annb_dataset *ds, int metric, int na, int32_t *A_dev,
                  double *Dam, DevBuf &scratch);
int transpose_D(annb_ctx *c, const double *Dam, int64_t n, int na, double *Dpm);
int pair_dists_f32_perm(annb_ctx *c, const annb_dataset *ds, int metric, const int32_t *I,
                        const int32_t *J, const int32_t *perm, int64_t n, float *out);

// ---------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------

// D (na, n) float64 anchor-major -> float32 (na, npad) + per-point metadata
__global__ void build_meta_kernel(const double *__restrict__ D64, int64_t n, int64_t npad, int na,
                                  int locality, int loc_thresh, float *__restrict__ D32,
                                  float *__restrict__ Dpm, int dpitch, PointMeta *__restrict__ meta)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < npad;
         j += (int64_t)gridDim.x * blockDim.x) {
        PointMeta m;
        m.amask = 0;
        m.cA = 0;
        m.loc_t = (int8_t)loc_thresh;
        m.slot = -1;
        m.pad = 0;
        if (j < n) {
            double best = INFINITY;
            for (int a = 0; a < na; ++a) {
                const double v = D64[(int64_t)a * n + j];
                D32[(int64_t)a * npad + j] = (float)v;
                Dpm[j * dpitch + a] = (float)v;
                if (v < best) {  // first minimum (np.argmin)
                    best = v;
                    m.cA = a;
                }
            }
            // `locality` nearest anchors, ties to the lower anchor index (stable argsort, annchor.py:235)
            for (int l = 0; l < locality && l < na; ++l) {
                double bv = INFINITY;
                int ba = -1;
                for (int a = 0; a < na; ++a) {
                    if ((m.amask >> a) & 1ull) continue;
                    const double v = D64[(int64_t)a * n + j];
                    if (ba < 0 || v < bv) {
                        bv = v;
                        ba = a;
                    }
                }
                m.amask |= 1ull << ba;
            }
        } else {
            for (int a = 0; a < na; ++a) {
                D32[(int64_t)a * npad + j] = 0.0f;
                Dpm[j * dpitch + a] = 0.0f;
            }
            m.loc_t = 127;  // padding rows are never candidates
        }
        meta[j] = m;
    }
}

__global__ void set_slots_kernel(const int32_t *__restrict__ A, int nA, PointMeta *__restrict__ meta)
{
    // serial so that duplicated anchors resolve like the reference's loop (last one wins)
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int k = 0; k < nA; ++k) meta[A[k]].slot = (int8_t)k;
}

// per row: histogram of min(shared-anchor count, loc_thresh) over all points (get_check,
// utils.py:470-473), or (pass 2) the number of candidates under the relaxed thresholds
__global__ void __launch_bounds__(256)
locality_count_kernel(const PointMeta *__restrict__ meta, int64_t n, int loc_thresh, int pass,
                      int32_t *__restrict__ out /* pass 1: [n][8]; pass 2: [n] */, int rank, int world)
{
    __shared__ uint64_t s_mask[1024];
    __shared__ int8_t s_t[1024];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = i < n;
    if ((int)(blockIdx.x % world) != rank) {
        // another rank counts this block of rows: leave zeros for the sum all-reduce
        if (live) {
            if (pass == 1) {
#pragma unroll
                for (int q = 0; q < 8; ++q) out[i * 8 + q] = 0;
            } else {
                out[i] = 0;
            }
        }
        return;
    }
    const PointMeta mi = live ? meta[i] : PointMeta{0, 0, 127, -1, 0};
    int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t j0 = 0; j0 < n; j0 += 1024) {
        __syncthreads();
        for (int k = threadIdx.x; k < 1024; k += blockDim.x) {
            const int64_t j = j0 + k;
            s_mask[k] = j < n ? meta[j].amask : 0ull;
            s_t[k] = j < n ? meta[j].loc_t : 127;
        }
        __syncthreads();
        const int lim = (int)min((int64_t)1024, n - j0);
        for (int k = 0; k < lim; ++k) {
            const int c = __popcll(mi.amask & s_mask[k]);
            if (pass == 1) {
                if (loc_thresh <= 1) {
                    cnt[1] += (c >= 1);  // default loc_thresh = 1: one counter; cnt[0] follows from n
                } else {
                    const int cc = c < loc_thresh ? c : loc_thresh;
#pragma unroll
                    for (int q = 0; q < 8; ++q) cnt[q] += (cc == q);
                }
            } else {
                const int t = mi.loc_t < s_t[k] ? mi.loc_t : s_t[k];
                cnt[0] += (c >= t) && (j0 + k != i);
            }
        }
    }
    if (!live) return;
    if (pass == 1) {
        if (loc_thresh <= 1) cnt[0] = (int)n - cnt[1];  // (loc_thresh = 0: everything is in bin 0 = cnt[0] + cnt[1])
        if (loc_thresh == 0) {
            cnt[0] = (int)n;
            cnt[1] = 0;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) out[i * 8 + q] = cnt[q];
    } else {
        out[i] = cnt[0];
    }
}

__global__ void set_loc_t_kernel(PointMeta *__restrict__ meta, const int8_t *__restrict__ t, int64_t n)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x)
        meta[j].loc_t = t[j];
}

__global__ void fill_u64_kernel(uint64_t *p, int64_t n, uint64_t v)
{
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x)
        p[j] = v;
}

// insert / overwrite entries of the known-pair hash map and raise their flag bits
__global__ void hash_insert_kernel(HashSlot *__restrict__ htab, uint64_t hmask,
                                   const int32_t *__restrict__ I, const int32_t *__restrict__ J,
                                   const float *__restrict__ a, const float *__restrict__ b,
                                   uint32_t kind, int64_t m)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m;
         p += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)I[p], j = (uint32_t)J[p];
        if (i == j) continue;
        const uint32_t lo = i < j ? i : j, hi = i < j ? j : i;
        const uint64_t key = pair_key(lo, hi);
        const uint64_t stored = key | kind_bits(kind);
        uint64_t h = mix64(key) & hmask;
        for (;;) {
            unsigned long long *kp = reinterpret_cast<unsigned long long *>(&htab[h].key);
            const uint64_t prev = atomicCAS(kp, HKEY_EMPTY, stored);
            if (prev == HKEY_EMPTY) {
                htab[h].a = a ? a[p] : 0.0f;
                htab[h].b = b ? b[p] : 0.0f;
                break;
            }
            if ((prev & HKEY_MASK) == key) {
                if (kind != KIND_KNOWN && kind_of(prev) == KIND_KNOWN) break;  // never downgrade
                atomicExch(kp, stored);
                htab[h].a = a ? a[p] : 0.0f;
                htab[h].b = b ? b[p] : 0.0f;
                break;
            }
            h = (h + 1) & hmask;
        }
    }
}

// ---- per-tile lists of the store's entries (what the sweeps read; see sweep.cuh) ----------------
__global__ void tl_count_kernel(const HashSlot *__restrict__ htab, uint64_t cap, int T, int32_t *__restrict__ cnt)
{
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap;
         s += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = htab[s].key;
        if (k == HKEY_EMPTY || kind_of(k) == KIND_NONE) continue;
        const uint64_t key = k & HKEY_MASK;
        const uint32_t lo = (uint32_t)(key >> 32), hi = (uint32_t)(key & 0xffffffffu);
        atomicAdd(&cnt[tile_index((int)(lo >> 7), (int)(hi >> 7), T)], 1);
    }
}
__global__ void tl_fill_kernel(const HashSlot *__restrict__ htab, uint64_t cap, int T,
                               const long long *__restrict__ ptr, int32_t *__restrict__ cursor,
                               uint32_t *__restrict__ code, float *__restrict__ a, float *__restrict__ b)
{
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap;
         s += (uint64_t)gridDim.x * blockDim.x) {
        const HashSlot e = htab[s];
        if (e.key == HKEY_EMPTY || kind_of(e.key) == KIND_NONE) continue;
        const uint64_t key = e.key & HKEY_MASK;
        const uint32_t lo = (uint32_t)(key >> 32), hi = (uint32_t)(key & 0xffffffffu);
        const int64_t t = tile_index((int)(lo >> 7), (int)(hi >> 7), T);
        const long long p = ptr[t] + atomicAdd(&cursor[t], 1);
        code[p] = ((lo & 127u) << 9) | ((hi & 127u) << 2) | kind_of(e.key);
        a[p] = e.a;
        b[p] = e.b;
    }
}

__global__ void hash_rehash_kernel(const HashSlot *__restrict__ otab, uint64_t ocap,
                                   HashSlot *__restrict__ htab, uint64_t hmask)
{
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < ocap;
         s += (uint64_t)gridDim.x * blockDim.x) {
        const HashSlot e = otab[s];
        if (e.key == HKEY_EMPTY) continue;
        uint64_t h = mix64(e.key & HKEY_MASK) & hmask;
        for (;;) {
            const uint64_t prev = atomicCAS(reinterpret_cast<unsigned long long *>(&htab[h].key),
                                            HKEY_EMPTY, e.key);
            if (prev == HKEY_EMPTY) {
                htab[h].a = e.a;
                htab[h].b = e.b;
                break;
            }
            h = (h + 1) & hmask;
        }
    }
}

// retire the transient FORCED marks after the iteration-0 selection (annchor.py:374-379 overwrites
// every not-computed RefineApprox on the next iteration)
__global__ void hash_retire_forced_kernel(HashSlot *__restrict__ htab, uint64_t cap)
{
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap;
         s += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = htab[s].key;
        if (k != HKEY_EMPTY && kind_of(k) == KIND_FORCED) htab[s].key = k & HKEY_MASK;  // KIND_NONE
    }
}

// number of anchors that are candidates of each point (those pairs are "computed" from the start,
// annchor.py:286-301)
__global__ void anchor_cand_kernel(const PointMeta *__restrict__ meta, int64_t n,
                                   const int32_t *__restrict__ A, int nA, int32_t *__restrict__ out)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const PointMeta mi = meta[i];
        int cnt = 0;
        for (int k = 0; k < nA; ++k) {
            const int a = A[k];
            const PointMeta ma = meta[a];
            if (a != i && ma.slot == k && __popcll(mi.amask & ma.amask) >= (mi.loc_t < ma.loc_t ? mi.loc_t : ma.loc_t))
                ++cnt;
        }
        out[i] = cnt;
    }
}

__global__ void split_ij32_kernel(const int64_t *__restrict__ ij, const double *__restrict__ d,
                                  int64_t n, int32_t *__restrict__ I, int32_t *__restrict__ J,
                                  float *__restrict__ df)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n;
         p += (int64_t)gridDim.x * blockDim.x) {
        I[p] = (int32_t)ij[2 * p];
        J[p] = (int32_t)ij[2 * p + 1];
        if (d) df[p] = (float)d[p];
    }
}

// features [lb, ub, dad] of explicit pairs in the sweep's float32 arithmetic (sampler output)
__global__ void pair_features_kernel(View V, const int32_t *__restrict__ I,
                                     const int32_t *__restrict__ J, int64_t m,
                                     float *__restrict__ feat /* (m,3) */)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int i = I[p], j = J[p];
        float lb = 0.0f, ub = INFINITY;
        for (int a = 0; a < V.na; ++a) {
            const float x = V.D32[(int64_t)a * V.npad + i], y = V.D32[(int64_t)a * V.npad + j];
            lb = fmaxf(lb, fabsf(x - y));
            ub = fminf(ub, x + y);
        }
        const PointMeta mi = V.meta[i], mj = V.meta[j];
        const float dad = 0.5f * (V.D32[(int64_t)mj.cA * V.npad + i] + V.D32[(int64_t)mi.cA * V.npad + j]);
        const uint32_t lo = i < j ? i : j, hi = i < j ? j : i;
        float ta = 0.0f, tb = 0.0f;
        if (hash_lookup(V, pair_key(lo, hi), ta, tb) == KIND_TIGHT) {
            lb = fmaxf(lb, ta);
            ub = fminf(ub, tb);
        }
        feat[3 * p] = lb;
        feat[3 * p + 1] = ub;
        feat[3 * p + 2] = dad;
    }
}

// store entry of explicit pairs: kind (0 none, 1 known, 2 tightened, 3 forced) and its two values
__global__ void pair_state_kernel(View V, const int32_t *__restrict__ I, const int32_t *__restrict__ J,
                                  int64_t m, int32_t *__restrict__ kind, float *__restrict__ ab /* (m,2) */)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m;
         p += (int64_t)gridDim.x * blockDim.x) {
        const int i = I[p], j = J[p];
        const uint32_t lo = i < j ? i : j, hi = i < j ? j : i;
        float a = 0.0f, b = 0.0f;
        kind[p] = i == j ? 0 : (int32_t)hash_lookup(V, pair_key(lo, hi), a, b);
        ab[2 * p] = a;
        ab[2 * p + 1] = b;
    }
}

// ---- selection over the emitted list -----------------------------------------------------------
// block-wide exclusive scan of a packed pair of 16-bit counts (lo | hi << 16); returns this
// thread's exclusive prefix and the block total (256 threads, each count <= 8 per half)
__device__ __forceinline__ uint32_t block_scan_packed(uint32_t v, uint32_t *s_warp /* [9] */, uint32_t &total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();  // protects s_warp across successive calls
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const uint32_t t = s_warp[q];
        if (q < w) base += t;
        tot += t;
    }
    total = tot;
    return base + x - v;
}

constexpr int PE_ITEMS = 8;  // entries per thread per block step

// selected: lvl > c1 or (lvl == c1 and mix <= thr1); next: not selected and (lvl > c2 or
// (lvl == c2 and mix <= thr2)).  One pair of global atomics per 2048 entries.
__global__ void __launch_bounds__(256)
partition_emitted_kernel(const uint64_t *__restrict__ keys, const uint16_t *__restrict__ lvl, int64_t E,
                         uint64_t salt, int c1, uint64_t thr1, int c2, uint64_t thr2, int32_t *__restrict__ sel_i,
                         int32_t *__restrict__ sel_j, int32_t *__restrict__ nxt_i,
                         int32_t *__restrict__ nxt_j, unsigned long long *__restrict__ cnt /* [0] sel, [1] next */,
                         int64_t cap_sel, int64_t cap_next)
{
    __shared__ uint32_t s_warp[9];
    __shared__ unsigned long long s_base[2];
    const int64_t step = 256 * PE_ITEMS;
    for (int64_t b0 = blockIdx.x * step; b0 < E; b0 += (int64_t)gridDim.x * step) {
        uint64_t key[PE_ITEMS];
        uint8_t cls[PE_ITEMS];  // 1 = selected, 2 = next
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < PE_ITEMS; ++k) {
            const int64_t p = b0 + k * 256 + threadIdx.x;
            cls[k] = 0;
            key[k] = 0;
            if (p < E) {
                const int l = lvl[p];
                key[k] = keys[p];
                const uint64_t mk = tie_key((uint32_t)(key[k] >> 32), (uint32_t)key[k], salt);
                const bool sel = l > c1 || (l == c1 && mk <= thr1);
                const bool nxt = !sel && (l > c2 || (l == c2 && mk <= thr2));
                cls[k] = sel ? 1 : (nxt ? 2 : 0);
                mine += sel ? 1u : (nxt ? 0x10000u : 0u);
            }
        }
        uint32_t total;
        const uint32_t ex = block_scan_packed(mine, s_warp, total);
        if (threadIdx.x == 0) {
            s_base[0] = (total & 0xffff) ? atomicAdd(&cnt[0], (unsigned long long)(total & 0xffff)) : 0ull;
            s_base[1] = (total >> 16) ? atomicAdd(&cnt[1], (unsigned long long)(total >> 16)) : 0ull;
        }
        __syncthreads();
        unsigned long long ps = s_base[0] + (ex & 0xffff), pn = s_base[1] + (ex >> 16);
#pragma unroll
        for (int k = 0; k < PE_ITEMS; ++k) {
            if (cls[k] == 1) {
                if ((int64_t)ps < cap_sel) {
                    sel_i[ps] = (int32_t)(key[k] >> 32);
                    sel_j[ps] = (int32_t)(key[k] & 0xffffffffu);
                }
                ++ps;
            } else if (cls[k] == 2) {
                if ((int64_t)pn < cap_next) {
                    nxt_i[pn] = (int32_t)(key[k] >> 32);
                    nxt_j[pn] = (int32_t)(key[k] & 0xffffffffu);
                }
                ++pn;
            }
        }
    }
}

// ---- CSR of exactly-known pairs ---------------------------------------------------------------
__global__ void known_degree_kernel(const HashSlot *__restrict__ htab, uint64_t cap,
                                    int32_t *__restrict__ deg)
{
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap;
         s += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t k = htab[s].key;
        if (k == HKEY_EMPTY || kind_of(k) != KIND_KNOWN) continue;
        const uint64_t key = k & HKEY_MASK;
        atomicAdd(&deg[(uint32_t)(key >> 32)], 1);
        atomicAdd(&deg[(uint32_t)(key & 0xffffffffu)], 1);
    }
}

// exclusive scan int32 -> int64 (out[n] = total) in three passes: per-block sums, scan of the sums by
// one block, per-block scan with offsets.  n ranges from N (CSR row pointers) to N^2 / 32768 (tile lists).
constexpr int SCAN_CHUNK = 1024 * 8;

__device__ __forceinline__ int64_t block_exclusive_scan_1024(int64_t v, int64_t *s_warp /* [33] */, int64_t &total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();
    if (lane == 31) s_warp[w] = x;
    __syncthreads();
    if (w == 0) {
        int64_t t = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += y;
        }
        s_warp[lane] = t;
    }
    __syncthreads();
    total = s_warp[31];
    return x - v + (w > 0 ? s_warp[w - 1] : 0);
}

__global__ void __launch_bounds__(1024) scan_block_sums_kernel(const int32_t *__restrict__ in, int64_t n,
                                                               int64_t *__restrict__ sums)
{
    __shared__ int64_t s_warp[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * 8;
    int64_t v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (base + k < n) v += in[base + k];
    int64_t total;
    block_exclusive_scan_1024(v, s_warp, total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_sums_kernel(int64_t *__restrict__ sums, int64_t nb)
{
    __shared__ int64_t s_warp[33];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += 1024) {
        const int64_t idx = base + threadIdx.x;
        const int64_t v = idx < nb ? sums[idx] : 0;
        int64_t total;
        const int64_t ex = block_exclusive_scan_1024(v, s_warp, total);
        const int64_t carry = s_carry;
        if (idx < nb) sums[idx] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[nb] = s_carry;
}

__global__ void __launch_bounds__(1024) scan_apply_kernel(const int32_t *__restrict__ in, int64_t n,
                                                          const int64_t *__restrict__ sums, int64_t nb,
                                                          int64_t *__restrict__ out)
{
    __shared__ int64_t s_warp[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)threadIdx.x * 8;
    int32_t x[8];
    int64_t v = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        x[k] = base + k < n ? in[base + k] : 0;
        v += x[k];
    }
    int64_t total;
    int64_t run = sums[blockIdx.x] + block_exclusive_scan_1024(v, s_warp, total);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (base + k < n) out[base + k] = run;
        run += x[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = sums[nb];
}

__global__ void known_fill_kernel(const HashSlot *__restrict__ htab, uint64_t cap,
                                  const int64_t *__restrict__ ptr, int32_t *__restrict__ cursor,
                                  int32_t *__restrict__ ids, float *__restrict__ ds)
{
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap;
         s += (uint64_t)gridDim.x * blockDim.x) {
        const HashSlot e = htab[s];
        if (e.key == HKEY_EMPTY || kind_of(e.key) != KIND_KNOWN) continue;
        const uint64_t key = e.key & HKEY_MASK;
        const uint32_t i = (uint32_t)(key >> 32), j = (uint32_t)(key & 0xffffffffu);
        const float d = e.a;
        int64_t p = ptr[i] + atomicAdd(&cursor[i], 1);
        ids[p] = (int32_t)j;
        ds[p] = d;
        p = ptr[j] + atomicAdd(&cursor[j], 1);
        ids[p] = (int32_t)i;
        ds[p] = d;
    }
}

// ---- update_bounds (utils.py:304-352) for the look-ahead pairs -------------------------------
// Every point k with both d(i,k) and d(j,k) known acts as an extra anchor for the pair (i,j):
//   lb = max_k |d_ik - d_jk|,  ub = min_k (d_ik + d_jk)   over k in N(i) & N(j).
// The reference merge-joins two id-sorted lists per pair.  Here the pairs are grouped by their
// lower endpoint i; one CTA owns an i, puts N(i) into a shared-memory hash table once, and each
// of its warps streams the (coalesced) list of one partner j at a time, probing the table.
// group key of a pair = its endpoint with the longer known-neighbour list (ties: lower id): that
// list is hashed once per group, the shorter partner lists are streamed
__device__ __forceinline__ void group_key(const int64_t *__restrict__ kptr, int a, int b, int &key, int &other)
{
    const int64_t da = kptr[a + 1] - kptr[a], db = kptr[b + 1] - kptr[b];
    const bool a_key = da > db || (da == db && a < b);
    key = a_key ? a : b;
    other = a_key ? b : a;
}

// group sizes, and per key row the list entries its group will stream (scheduling weight)
__global__ void count_by_lo_kernel(const int32_t *__restrict__ I, const int32_t *__restrict__ J, int64_t m,
                                   const int64_t *__restrict__ kptr, int32_t *__restrict__ cnt,
                                   unsigned long long *__restrict__ work, unsigned long long *__restrict__ total)
{
    unsigned long long mine = 0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m;
         p += (int64_t)gridDim.x * blockDim.x) {
        int key, other;
        group_key(kptr, I[p], J[p], key, other);
        atomicAdd(&cnt[key], 1);
        const unsigned long long d = (unsigned long long)(kptr[other + 1] - kptr[other]);
        atomicAdd(&work[key], d);
        mine += d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, mine);
}

__global__ void scatter_by_lo_kernel(const int32_t *__restrict__ I, const int32_t *__restrict__ J, int64_t m,
                                     const int64_t *__restrict__ kptr, const int64_t *__restrict__ ptr,
                                     int32_t *__restrict__ cursor, int32_t *__restrict__ gJ /* partner */,
                                     int32_t *__restrict__ gsrc)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m;
         p += (int64_t)gridDim.x * blockDim.x) {
        int key, other;
        group_key(kptr, I[p], J[p], key, other);
        const int64_t q = ptr[key] + atomicAdd(&cursor[key], 1);
        gJ[q] = other;
        gsrc[q] = (int32_t)p;
    }
}

// rows ordered by (closest anchor, distance to it): a counting sort over 64 x 1024 buckets.  Points that
// are close to each other get close positions, so the CTAs that run at the same time work on rows whose
// partner lists overlap and the streamed lists hit in L2.
constexpr int ROW_BUCKETS = kMaxAnchors * 1024;
__device__ __forceinline__ int row_bucket(const PointMeta &m, const float *__restrict__ D32, int64_t npad, int64_t i,
                                          float inv_scale)
{
    const float d = D32[(int64_t)m.cA * npad + i];
    int q = (int)(d * inv_scale);
    q = q < 0 ? 0 : (q > 1023 ? 1023 : q);
    return m.cA * 1024 + q;
}
__global__ void row_bucket_hist_kernel(const PointMeta *__restrict__ meta, const float *__restrict__ D32,
                                       int64_t npad, int64_t n, float inv_scale, int32_t *__restrict__ hist)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[row_bucket(meta[i], D32, npad, i, inv_scale)], 1);
}
__global__ void row_bucket_scatter_kernel(const PointMeta *__restrict__ meta, const float *__restrict__ D32,
                                          int64_t npad, int64_t n, float inv_scale,
                                          const int64_t *__restrict__ ptr, int32_t *__restrict__ cursor,
                                          int32_t *__restrict__ order)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int b = row_bucket(meta[i], D32, npad, i, inv_scale);
        order[ptr[b] + atomicAdd(&cursor[b], 1)] = (int32_t)i;
    }
}

__global__ void tighten_work_kernel(const int64_t *__restrict__ kptr, const int64_t *__restrict__ gptr,
                                    const int32_t *__restrict__ gJ, int64_t n,
                                    unsigned long long *__restrict__ out /* [0] sum deg_j, [1] sum min, [2] max deg */,
                                    unsigned long long *__restrict__ row_work /* [n] sum deg_j of the row's group */)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long di = kptr[i + 1] - kptr[i];
        unsigned long long s = 0, sm = 0;
        for (int64_t g = gptr[i]; g < gptr[i + 1]; ++g) {
            const int j = gJ[g];
            const unsigned long long dj = kptr[j + 1] - kptr[j];
            s += dj;
            sm += dj < di ? dj : di;
        }
        if (row_work) row_work[i] = s;
        atomicAdd(&out[0], s);
        atomicAdd(&out[1], sm);
        atomicMax(&out[2], di);
    }
}

constexpr int TG_SLOTS = 8192;   // shared-memory table slots: 2048 buckets of 4 ids (+ 4 values), 64 KB -> 3 CTAs / SM
                                 // (a 4096-slot table at 6 CTAs / SM is 2.3x slower: hub rows then need several
                                 // chunks and every chunk re-streams all partner lists)
constexpr int TG_CHUNK = 2048;   // entries of N(i) hashed at a time (<= 1 key per 3-slot bucket)

// rows whose group streams more than `thr` list entries are queued first (longest-processing-time
// first keeps the dynamic row scheduler's tail short)
__global__ void tighten_heavy_kernel(const unsigned long long *__restrict__ work, int64_t n,
                                     unsigned long long thr, int32_t *__restrict__ heavy,
                                     unsigned long long *__restrict__ n_heavy)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x)
        if (work[i] >= thr) heavy[atomicAdd(n_heavy, 1ull)] = (int32_t)i;
}

// K2b.  Dynamic row scheduling (one atomic per row), bucketed shared-memory hash (4 ids per bucket,
// one 128-bit probe, no divergent chains at load <= 0.25), partner lists streamed with the next
// 128 entries in flight while the current ones are probed.
#ifdef ANNB_VARIANT_TGU8
constexpr int TGU = 8;  // 32-entry loads of a partner list in flight per warp
#else
constexpr int TGU = 4;
#endif

__global__ void __launch_bounds__(512)
tighten_grouped_kernel(View V, const int64_t *__restrict__ kptr, const int32_t *__restrict__ kids,
                       const float *__restrict__ kds, const int64_t *__restrict__ gptr,
                       const int32_t *__restrict__ gJ, const int32_t *__restrict__ gsrc,
                       const int32_t *__restrict__ row_order, const int32_t *__restrict__ heavy,
                       const unsigned long long *__restrict__ sched /* [0] n_heavy, [1] next ticket */,
                       const unsigned long long *__restrict__ work, unsigned long long heavy_thr,
                       int has_tight, float *__restrict__ out_lb, float *__restrict__ out_ub,
                       uint8_t *__restrict__ improved)
{
    extern __shared__ __align__(16) unsigned char tg_smem[];
    int32_t *h_id = reinterpret_cast<int32_t *>(tg_smem);
    float *h_d = reinterpret_cast<float *>(h_id + TG_SLOTS);
    __shared__ long long s_ticket;
    __shared__ int s_next;
    const int lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const long long n_heavy = (long long)sched[0];
    unsigned long long *ticket = const_cast<unsigned long long *>(sched) + 1;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_ticket = (long long)atomicAdd(ticket, 1ull);
        __syncthreads();
        const long long tk = s_ticket;
        if (tk >= n_heavy + V.n) break;
        int64_t i;
        if (tk < n_heavy) {
            i = heavy[tk];
        } else {
            // rows in closest-anchor order: CTAs running at the same time then work on points of the
            // same neighbourhood, whose partner lists overlap -> the streamed lists hit in L2
            i = row_order[tk - n_heavy];
            if (work[i] >= heavy_thr) continue;  // already done from the heavy queue
        }
        const int64_t g0 = gptr[i], g1 = gptr[i + 1];
        if (g0 == g1) continue;
        const int64_t bi = kptr[i];
        const int mi = (int)(kptr[i + 1] - bi);
        // this row's anchor distances, one anchor per lane (coalesced 128 B row of Dpm)
        float xa[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) xa[q] = (lane + 32 * q < V.na) ? __ldg(V.Dpm + i * V.dpitch + lane + 32 * q) : 0.0f;
        // N(i) goes through the shared-memory table in chunks (one chunk for all but hub points)
        for (int c0 = 0; c0 == 0 || c0 < mi; c0 += TG_CHUNK) {
            const int mc = min(TG_CHUNK, mi - c0);
            const bool first = c0 == 0, last = c0 + TG_CHUNK >= mi;
            // bucket = 4 words: 3 ids + an overflow flag (set when a key that hashes here had to be placed
            // further on), so a probe walks on only when it must -- a full bucket alone does not force it.
            // <= 0.5 keys per bucket where the table allows it (P[overflow] = 0.2 %), <= 1 otherwise.
            int nb = 16;
            while (nb < 2 * mc && nb < TG_SLOTS / 4) nb <<= 1;
            const uint32_t bmask = (uint32_t)nb - 1;
            __syncthreads();
            for (int k = threadIdx.x; k < nb * 4; k += blockDim.x) h_id[k] = (k & 3) == 3 ? 0 : -1;
            if (threadIdx.x == 0) s_next = 0;
            __syncthreads();
            for (int k = threadIdx.x; k < mc; k += blockDim.x) {
                const int32_t id = kids[bi + c0 + k];
                uint32_t bk = ((uint32_t)id * 2654435761u >> 7) & bmask;
                for (;;) {
                    int sl = 0;
                    for (; sl < 3; ++sl)
                        if (atomicCAS(&h_id[bk * 4 + sl], -1, id) == -1) break;
                    if (sl < 3) {
                        h_d[bk * 4 + sl] = kds[bi + c0 + k];
                        break;
                    }
                    h_id[bk * 4 + 3] = 1;  // overflowed: probes for absent ids must walk on from here
                    bk = (bk + 1) & bmask;
                }
            }
            __syncthreads();
            for (;;) {
                int gq = 0;
                if (lane == 0) gq = atomicAdd(&s_next, 1);
                gq = __shfl_sync(0xffffffffu, gq, 0);
                const int64_t g = g0 + gq;
                if (g >= g1) break;
                const int j = gJ[g];
                const int64_t bj = kptr[j];
                const int mj = (int)(kptr[j + 1] - bj);
                float lb = 0.0f, ub = INFINITY;
                int32_t idv[TGU], idn[TGU];
                float yv[TGU], yn[TGU];
#pragma unroll
                for (int u = 0; u < TGU; ++u) {
                    const int k = u * 32 + lane;
                    idv[u] = k < mj ? __ldg(kids + bj + k) : -2;
                    yv[u] = k < mj ? __ldg(kds + bj + k) : 0.0f;
                }
                for (int k0 = 0; k0 < mj; k0 += 32 * TGU) {
                    if (k0 + 32 * TGU < mj) {
#pragma unroll
                        for (int u = 0; u < TGU; ++u) {
                            const int k = k0 + 32 * TGU + u * 32 + lane;
                            idn[u] = k < mj ? __ldg(kids + bj + k) : -2;
                            yn[u] = k < mj ? __ldg(kds + bj + k) : 0.0f;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < TGU; ++u) {
                        // branch-free probe of the home bucket: one 128-bit load (3 ids + overflow flag),
                        // one load of the candidate value; padding lanes carry id = -2 and match nothing
                        const int32_t id = idv[u];
                        uint32_t bk = ((uint32_t)id * 2654435761u >> 7) & bmask;
                        int4 t = *reinterpret_cast<const int4 *>(h_id + bk * 4);
                        float x = h_d[bk * 4 + (t.y == id ? 1 : (t.z == id ? 2 : 0))];
                        bool hit = (t.x == id) | (t.y == id) | (t.z == id);
                        bool more = !hit && t.w != 0 && id >= 0;  // rare: the home bucket overflowed
                        while (__any_sync(0xffffffffu, more)) {
                            if (more) {
                                bk = (bk + 1) & bmask;
                                t = *reinterpret_cast<const int4 *>(h_id + bk * 4);
                                x = h_d[bk * 4 + (t.y == id ? 1 : (t.z == id ? 2 : 0))];
                                hit = (t.x == id) | (t.y == id) | (t.z == id);
                                more = !hit && t.w != 0;
                            }
                        }
                        lb = fmaxf(lb, hit ? fabsf(x - yv[u]) : 0.0f);
                        ub = fminf(ub, hit ? x + yv[u] : INFINITY);
                    }
#pragma unroll
                    for (int u = 0; u < TGU; ++u) {
                        idv[u] = idn[u];
                        yv[u] = yn[u];
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    lb = fmaxf(lb, __shfl_xor_sync(0xffffffffu, lb, o));
                    ub = fminf(ub, __shfl_xor_sync(0xffffffffu, ub, o));
                }
                const int32_t p = gsrc[g];
                if (!first) {  // combine with the earlier chunks of N(i)
                    lb = fmaxf(lb, out_lb[p]);
                    ub = fminf(ub, out_ub[p]);
                }
                if (!last) {
                    if (lane == 0) {
                        out_lb[p] = lb;
                        out_ub[p] = ub;
                    }
                    continue;
                }
                // anchor bounds + any earlier tightening of this pair: lanes over anchors
                float l0 = 0.0f, u0 = INFINITY;
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (lane + 32 * q < V.na) {
                        const float y = __ldg(V.Dpm + (int64_t)j * V.dpitch + lane + 32 * q);
                        l0 = fmaxf(l0, fabsf(xa[q] - y));
                        u0 = fminf(u0, xa[q] + y);
                    }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    l0 = fmaxf(l0, __shfl_xor_sync(0xffffffffu, l0, o));
                    u0 = fminf(u0, __shfl_xor_sync(0xffffffffu, u0, o));
                }
                if (lane == 0) {
                    if (has_tight) {
                        float ta = 0.0f, tb = 0.0f;
                        const uint32_t plo = (uint32_t)min((int64_t)j, i), phi = (uint32_t)max((int64_t)j, i);
                        if (hash_lookup(V, pair_key(plo, phi), ta, tb) == KIND_TIGHT) {
                            l0 = fmaxf(l0, ta);
                            u0 = fminf(u0, tb);
                        }
                    }
                    out_lb[p] = fmaxf(lb, l0);
                    out_ub[p] = fminf(ub, u0);
                    improved[p] = (lb > l0 || ub < u0) ? 1 : 0;
                }
            }
        }
    }
}

// K2b, bitmap variant (the default while N bits fit shared memory).  The key row's known-neighbour set
// N(i) is a bit per point id in shared memory plus a per-word prefix count, so membership of a streamed
// id is ONE shared load and a bit test, and a hit finds its distance by bit rank (prefix + popcount) in a
// dense value array -- ~25 issue slots per 32 streamed entries against ~100 for the bucket-hash probe
// above.  Bits are set and cleared per row from the row's own list (no full clears); rows with more than
// TB_CHUNK known distances go through in chunks whose partial bounds are combined.
constexpr int TB_CHUNK = 4096;

template <int NT_>
__global__ void __launch_bounds__(NT_)
tighten_bitmap_kernel(View V, const int64_t *__restrict__ kptr, const int32_t *__restrict__ kids,
                      const float *__restrict__ kds, const int64_t *__restrict__ gptr,
                      const int32_t *__restrict__ gJ, const int32_t *__restrict__ gsrc,
                      const int32_t *__restrict__ row_order, const int32_t *__restrict__ heavy,
                      const unsigned long long *__restrict__ sched /* [0] n_heavy, [1] next ticket */,
                      const unsigned long long *__restrict__ work, unsigned long long heavy_thr,
                      int has_tight, float *__restrict__ out_lb, float *__restrict__ out_ub,
                      uint8_t *__restrict__ improved, int W /* bitmap words, multiple of 32 */,
                      int chunk /* entries of N(i) per pass, <= TB_CHUNK */)
{
    extern __shared__ __align__(16) unsigned char tb_smem[];
    uint32_t *bits = reinterpret_cast<uint32_t *>(tb_smem);
    float *vals = reinterpret_cast<float *>(bits + W);
    uint16_t *pre = reinterpret_cast<uint16_t *>(vals + TB_CHUNK);
    __shared__ long long s_ticket;
    __shared__ int s_next;
    __shared__ int s_wtot[NT_ / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT_ / 32;
    const int S = ((W + NW - 1) / NW + 31) / 32 * 32;  // words per warp segment of the prefix pass
    for (int k = tid; k < W; k += NT_) bits[k] = 0;
    const long long n_heavy = (long long)sched[0];
    unsigned long long *ticket = const_cast<unsigned long long *>(sched) + 1;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_ticket = (long long)atomicAdd(ticket, 1ull);
        __syncthreads();
        const long long tk = s_ticket;
        if (tk >= n_heavy + V.n) break;
        int64_t i;
        if (tk < n_heavy) {
            i = heavy[tk];
        } else {
            i = row_order[tk - n_heavy];
            if (work[i] >= heavy_thr) continue;  // already done from the heavy queue
        }
        const int64_t g0 = gptr[i], g1 = gptr[i + 1];
        if (g0 == g1) continue;
        const int64_t bi = kptr[i];
        const int mi = (int)(kptr[i + 1] - bi);
        float xa[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) xa[q] = (lane + 32 * q < V.na) ? __ldg(V.Dpm + i * V.dpitch + lane + 32 * q) : 0.0f;
        for (int c0 = 0; c0 == 0 || c0 < mi; c0 += chunk) {
            const int mc = min(chunk, mi - c0);
            const bool first = c0 == 0, last = c0 + chunk >= mi;
            // 1. membership bits of this chunk of N(i)
            for (int k = tid; k < mc; k += NT_) {
                const uint32_t id = (uint32_t)kids[bi + c0 + k];
                atomicOr(&bits[id >> 5], 1u << (id & 31));
            }
            if (tid == 0) s_next = 0;
            __syncthreads();
            // 2. pre[w] = set bits in the words before w: each warp scans its own segment, then adds
            //    the totals of the segments before it
            {
                const int w0 = warp * S, w1 = min(W, w0 + S);
                int run = 0;
                for (int b = w0; b < w1; b += 32) {
                    const int w = b + lane;
                    const int c = w < w1 ? __popc(bits[w]) : 0;
                    int x = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int y = __shfl_up_sync(0xffffffffu, x, o);
                        if (lane >= o) x += y;
                    }
                    if (w < w1) pre[w] = (uint16_t)(run + x - c);
                    run += __shfl_sync(0xffffffffu, x, 31);
                }
                if (lane == 0) s_wtot[warp] = run;
                __syncthreads();
                int off = 0;
                for (int q = 0; q < warp; ++q) off += s_wtot[q];
                if (off)
                    for (int w = w0 + lane; w < w1; w += 32) pre[w] = (uint16_t)(pre[w] + off);
            }
            __syncthreads();
            // 3. distances by bit rank
            for (int k = tid; k < mc; k += NT_) {
                const uint32_t id = (uint32_t)kids[bi + c0 + k];
                const uint32_t w = id >> 5;
                vals[pre[w] + __popc(bits[w] & ((1u << (id & 31)) - 1u))] = kds[bi + c0 + k];
            }
            __syncthreads();
            // 4. every warp streams the lists of the row's partners
            for (;;) {
                int gq = 0;
                if (lane == 0) gq = atomicAdd(&s_next, 1);
                gq = __shfl_sync(0xffffffffu, gq, 0);
                const int64_t g = g0 + gq;
                if (g >= g1) break;
                const int j = gJ[g];
                const int64_t bj = kptr[j];
                const int mj = (int)(kptr[j + 1] - bj);
                float lb = 0.0f, ub = INFINITY;
                // ids are streamed (next 32 * TGU in flight); a partner's distance is only loaded when its
                // id hits the bitmap -- common neighbours are rare, so most entries cost one shared-memory
                // load and a bit test
                uint32_t idv[TGU], idn[TGU];
#pragma unroll
                for (int u = 0; u < TGU; ++u) {
                    const int k = u * 32 + lane;
                    idv[u] = k < mj ? (uint32_t)__ldg(kids + bj + k) : 0xffffffffu;
                }
                for (int k0 = 0; k0 < mj; k0 += 32 * TGU) {
                    if (k0 + 32 * TGU < mj) {
#pragma unroll
                        for (int u = 0; u < TGU; ++u) {
                            const int k = k0 + 32 * TGU + u * 32 + lane;
                            idn[u] = k < mj ? (uint32_t)__ldg(kids + bj + k) : 0xffffffffu;
                        }
                    }
                    uint32_t hits = 0;
                    int rk[TGU];
#pragma unroll
                    for (int u = 0; u < TGU; ++u) {
                        const uint32_t id = idv[u];
                        const uint32_t w = min(id >> 5, (uint32_t)(W - 1));  // padding lanes probe the last word
                        const uint32_t word = bits[w];
                        const uint32_t bit = 1u << (id & 31);
                        rk[u] = 0;
                        if ((word & bit) != 0u && id != 0xffffffffu) {
                            hits |= 1u << u;
                            rk[u] = pre[w] + __popc(word & (bit - 1u));
                        }
                    }
                    if (hits) {
                        float y[TGU];
#pragma unroll
                        for (int u = 0; u < TGU; ++u)
                            y[u] = ((hits >> u) & 1u) ? __ldg(kds + bj + k0 + u * 32 + lane) : 0.0f;
#pragma unroll
                        for (int u = 0; u < TGU; ++u)
                            if ((hits >> u) & 1u) {
                                const float x = vals[rk[u]];
                                lb = fmaxf(lb, fabsf(x - y[u]));
                                ub = fminf(ub, x + y[u]);
                            }
                    }
#pragma unroll
                    for (int u = 0; u < TGU; ++u) idv[u] = idn[u];
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    lb = fmaxf(lb, __shfl_xor_sync(0xffffffffu, lb, o));
                    ub = fminf(ub, __shfl_xor_sync(0xffffffffu, ub, o));
                }
                const int32_t p = gsrc[g];
                if (!first) {  // combine with the earlier chunks of N(i)
                    lb = fmaxf(lb, out_lb[p]);
                    ub = fminf(ub, out_ub[p]);
                }
                if (!last) {
                    if (lane == 0) {
                        out_lb[p] = lb;
                        out_ub[p] = ub;
                    }
                    continue;
                }
                // anchor bounds + any earlier tightening of this pair: lanes over anchors
                float l0 = 0.0f, u0 = INFINITY;
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (lane + 32 * q < V.na) {
                        const float y = __ldg(V.Dpm + (int64_t)j * V.dpitch + lane + 32 * q);
                        l0 = fmaxf(l0, fabsf(xa[q] - y));
                        u0 = fminf(u0, xa[q] + y);
                    }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    l0 = fmaxf(l0, __shfl_xor_sync(0xffffffffu, l0, o));
                    u0 = fminf(u0, __shfl_xor_sync(0xffffffffu, u0, o));
                }
                if (lane == 0) {
                    if (has_tight) {
                        float ta = 0.0f, tb = 0.0f;
                        const uint32_t plo = (uint32_t)min((int64_t)j, i), phi = (uint32_t)max((int64_t)j, i);
                        if (hash_lookup(V, pair_key(plo, phi), ta, tb) == KIND_TIGHT) {
                            l0 = fmaxf(l0, ta);
                            u0 = fminf(u0, tb);
                        }
                    }
                    out_lb[p] = fmaxf(lb, l0);
                    out_ub[p] = fminf(ub, u0);
                    improved[p] = (lb > l0 || ub < u0) ? 1 : 0;
                }
            }
            __syncthreads();
            // 5. clear this chunk's bits again (cheaper than zeroing W words per row)
            for (int k = tid; k < mc; k += NT_) bits[(uint32_t)kids[bi + c0 + k] >> 5] = 0;
            __syncthreads();
        }
    }
}

__global__ void compact_improved_kernel(const int32_t *__restrict__ I, const int32_t *__restrict__ J,
                                        const float *__restrict__ lb, const float *__restrict__ ub,
                                        const uint8_t *__restrict__ improved, int64_t m,
                                        int32_t *__restrict__ oI, int32_t *__restrict__ oJ,
                                        float *__restrict__ olb, float *__restrict__ oub,
                                        unsigned long long *__restrict__ cnt)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m;
         p += (int64_t)gridDim.x * blockDim.x) {
        if (!improved[p]) continue;
        const unsigned long long s = atomicAdd(cnt, 1ull);
        oI[s] = I[p];
        oJ[s] = J[p];
        olb[s] = lb[p];
        oub[s] = ub[p];
    }
}

// get_nn (utils.py:383-429) over the computed pairs of each row: CSR known entries + anchor
// distances.  One block per row; nn-1 extract-min rounds ordered by (distance, neighbour id).
__global__ void __launch_bounds__(256)
neighbor_graph_kernel(View V, const double *__restrict__ D64, const int32_t *__restrict__ A, int nA,
                      const int64_t *__restrict__ ptr, const int32_t *__restrict__ ids,
                      const float *__restrict__ ds, int64_t *__restrict__ out_idx,
                      double *__restrict__ out_d, int32_t *__restrict__ deficient)
{
    __shared__ double s_d[8];
    __shared__ int s_id[8];
    __shared__ double s_pd;
    __shared__ int s_pid;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nn = V.nn;
    for (int64_t row = blockIdx.x; row < V.n; row += gridDim.x) {
        const PointMeta mr = V.meta[row];
        const int64_t beg = ptr[row];
        const int mk = (int)(ptr[row + 1] - beg);
        if (threadIdx.x == 0) {
            s_pd = -INFINITY;
            s_pid = -1;
            out_idx[row * nn] = row;
            out_d[row * nn] = 0.0;
        }
        __syncthreads();
        for (int r = 0; r < nn - 1; ++r) {
            const double pd = s_pd;
            const int pid = s_pid;
            double bd = INFINITY;
            int bid = INT32_MAX;
            auto offer = [&](double d, int id) {
                const bool after = d > pd || (d == pd && id > pid);
                if (after && (d < bd || (d == bd && id < bid))) {
                    bd = d;
                    bid = id;
                }
            };
            for (int k = threadIdx.x; k < mk; k += blockDim.x) offer((double)ds[beg + k], ids[beg + k]);
            if (mr.slot >= 0) {
                // anchor row: every candidate pair is computed (annchor.py:288-289)
                for (int64_t j = threadIdx.x; j < V.n; j += blockDim.x)
                    if (j != row && is_candidate(mr, V.meta[j]))
                        offer(D64[(int64_t)mr.slot * V.n + j], (int)j);
            } else {
                for (int k = threadIdx.x; k < nA; k += blockDim.x) {
                    const int a = A[k];
                    if (V.meta[a].slot == k && a != row && is_candidate(mr, V.meta[a]))
                        offer(D64[(int64_t)k * V.n + row], a);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double yd = __shfl_xor_sync(0xffffffffu, bd, o);
                const int yi = __shfl_xor_sync(0xffffffffu, bid, o);
                if (yd < bd || (yd == bd && yi < bid)) {
                    bd = yd;
                    bid = yi;
                }
            }
            if (lane == 0) {
                s_d[w] = bd;
                s_id[w] = bid;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                for (int q = 1; q < 8; ++q)
                    if (s_d[q] < bd || (s_d[q] == bd && s_id[q] < bid)) {
                        bd = s_d[q];
                        bid = s_id[q];
                    }
                s_pd = bd;
                s_pid = bid;
                if (bid == INT32_MAX) {
                    out_idx[row * nn + 1 + r] = -1;
                    out_d[row * nn + 1 + r] = INFINITY;
                    atomicAdd(deficient, 1);
                    s_pd = INFINITY;
                } else {
                    out_idx[row * nn + 1 + r] = bid;
                    out_d[row * nn + 1 + r] = bd;
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace annb

using namespace annb;

namespace annb {

// mixed keys of the emitted entries at one level (ties at a selection cut); one global atomic per
// 2048 entries
__global__ void __launch_bounds__(256)
compact_level_kernel(const uint64_t *__restrict__ keys, const uint16_t *__restrict__ lvl, int64_t E,
                     uint64_t salt, int level, uint64_t *__restrict__ out, int64_t out_cap,
                     unsigned long long *__restrict__ cnt)
{
    __shared__ uint32_t s_warp[9];
    __shared__ unsigned long long s_base;
    const int64_t step = 256 * PE_ITEMS;
    for (int64_t b0 = blockIdx.x * step; b0 < E; b0 += (int64_t)gridDim.x * step) {
        bool hit[PE_ITEMS];
        uint32_t mine = 0;
#pragma unroll
        for (int k = 0; k < PE_ITEMS; ++k) {
            const int64_t p = b0 + k * 256 + threadIdx.x;
            hit[k] = p < E && lvl[p] == level;
            mine += hit[k] ? 1u : 0u;
        }
        uint32_t total;
        const uint32_t ex = block_scan_packed(mine, s_warp, total);
        if (threadIdx.x == 0) s_base = total ? atomicAdd(cnt, (unsigned long long)total) : 0ull;
        __syncthreads();
        unsigned long long pos = s_base + ex;
#pragma unroll
        for (int k = 0; k < PE_ITEMS; ++k)
            if (hit[k]) {
                if ((int64_t)pos < out_cap) {
                    const uint64_t kk = keys[b0 + k * 256 + threadIdx.x];
                    out[pos] = tie_key((uint32_t)(kk >> 32), (uint32_t)kk, salt);
                }
                ++pos;
            }
    }
}

// 16-bit digit histogram of 64-bit keys that share `prefix` above the digit
__global__ void digit_hist_kernel(const uint64_t *__restrict__ keys, int64_t m, uint64_t prefix, int shift,
                                  uint32_t *__restrict__ hist /* 65536 */)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < m;
         p += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[p];
        if (shift < 48 && (k >> (shift + 16)) != prefix) continue;
        atomicAdd(&hist[(k >> shift) & 0xffff], 1u);
    }
}

__global__ void max_f32_kernel(const float *__restrict__ x, int64_t n, unsigned int *__restrict__ out)
{
    float m = 0.0f;
    for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n;
         j += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, x[j]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));  // non-negative floats order as uints
}

// a rank's partial per-row lists as records for the cross-rank merge: k1 values (flagged computed:
// they only feed the thresh list), then the k2 not-computed (value, id) pairs
__global__ void pack_partial_lists_kernel(const float *__restrict__ l1last, const float *__restrict__ l2v,
                                          const int32_t *__restrict__ l2i, const float *__restrict__ l1all,
                                          int64_t n, int k1, int k2, uint2 *__restrict__ rec,
                                          int32_t *__restrict__ cnt)
{
    const int K = k1 + k2;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n * K;
         q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = q / K;
        const int k = (int)(q % K);
        uint2 r;
        if (k < k1) {
            r = make_uint2(__float_as_uint(l1all[row * k1 + k]), 0x80000000u);
        } else {
            const int32_t id = l2i[row * k2 + (k - k1)];
            const float v = id < 0 ? INFINITY : l2v[row * k2 + (k - k1)];
            r = make_uint2(__float_as_uint(v), (uint32_t)(id < 0 ? 0 : id) | 0x40000000u);  // L2-only
        }
        rec[q] = r;
        if (k == 0) cnt[row] = K;
    }
    (void)l1last;
}

}  // namespace annb

static int grid_for_n(const annb_ctx *c, int64_t n, int per_block = 256)
{
    int64_t g = (n + per_block - 1) / per_block;
    const int64_t cap = (int64_t)c->num_sms * 16;
    return (int)(g < 1 ? 1 : (g < cap ? g : cap));
}

static int hash_alloc(annb_index *ix, uint64_t cap)
{
    annb_ctx *c = ix->ctx;
    ANNB_TRY(ix->htab.ensure(cap * sizeof(HashSlot)));
    ANNB_CUDA(cudaMemsetAsync(ix->htab.p, 0xff, cap * sizeof(HashSlot), c->stream));  // key = HKEY_EMPTY
    ix->hcap = cap;
    return ANNB_OK;
}

// make room for `extra` more entries: load factor <= 0.75 (look-ups are almost always hits -- the
// flag bitmap filters the misses -- so the longer probe chains of a fuller table cost little, and a
// power-of-two table at <= 0.5 would need 69 GB instead of 34 GB at N=1M, p_work=1e-3)
static int hash_reserve(annb_index *ix, int64_t extra)
{
    annb_ctx *c = ix->ctx;
    const uint64_t need = (uint64_t)(ix->hcount_ub + extra) * 4 / 3 + 1024;
    if (need <= ix->hcap) return ANNB_OK;
    uint64_t cap = ix->hcap ? ix->hcap : 1024;
    while (cap < need) cap <<= 1;
    DevBuf otab = ix->htab;
    const uint64_t ocap = ix->hcap;
    ix->htab = DevBuf();
    ANNB_TRY(hash_alloc(ix, cap));
    if (ocap)
        ANNB_LAUNCH(hash_rehash_kernel, grid_for_n(c, (int64_t)ocap), 256, 0, c->stream,
                    otab.as<HashSlot>(), ocap, ix->htab.as<HashSlot>(), cap - 1);
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    otab.release();
    return ANNB_OK;
}

// insert device-resident (I, J, a, b) with the given kind
static int hash_insert(annb_index *ix, const int32_t *I, const int32_t *J, const float *a,
                       const float *b, uint32_t kind, int64_t m)
{
    if (m == 0) return ANNB_OK;
    annb_ctx *c = ix->ctx;
    ANNB_TRY(hash_reserve(ix, m));
    ANNB_LAUNCH(hash_insert_kernel, grid_for_n(c, m), 256, 0, c->stream, ix->htab.as<HashSlot>(),
                ix->hcap - 1, I, J, a, b, kind, m);
    ix->hcount_ub += m;
    ix->tl_dirty = true;
    return ANNB_OK;
}

namespace annb {
// neighbour graph of a renumbered index back to the caller's numbering: rows and neighbour ids
__global__ void unpermute_graph_kernel(const int64_t *__restrict__ idx, const double *__restrict__ dist, int64_t n,
                                       int nn, const int32_t *__restrict__ order, int64_t *__restrict__ oidx,
                                       double *__restrict__ odist)
{
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n * nn; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = q / nn, col = q % nn;
        const int64_t id = idx[q];
        const int64_t o = (int64_t)order[row] * nn + col;
        oidx[o] = id >= 0 ? order[id] : -1;
        odist[o] = dist[q];
    }
}

// per tile: min / max of every anchor's distance over the tile's (real) points, and the closest anchors present
__global__ void __launch_bounds__(64)
tile_bounds_kernel(const float *__restrict__ D32, const PointMeta *__restrict__ meta, int64_t n, int64_t npad, int na,
                   float *__restrict__ lo, float *__restrict__ hi, uint64_t *__restrict__ cm)
{
    const int t = blockIdx.x, a = threadIdx.x;
    const int64_t p0 = (int64_t)t * TILE;
    const int cnt = (int)min((int64_t)TILE, n - p0);
    float mn = INFINITY, mx = -INFINITY;
    if (a < na)
        for (int p = 0; p < cnt; ++p) {
            const float v = D32[(int64_t)a * npad + p0 + p];
            mn = fminf(mn, v);
            mx = fmaxf(mx, v);
        }
    lo[(int64_t)t * kMaxAnchors + a] = a < na ? mn : 0.0f;
    hi[(int64_t)t * kMaxAnchors + a] = a < na ? mx : 0.0f;
    unsigned long long m = 0;
    for (int p = a; p < cnt; p += 64) m |= 1ull << meta[p0 + p].cA;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
    __shared__ unsigned long long s_m[2];
    if ((a & 31) == 0) s_m[a >> 5] = m;
    __syncthreads();
    if (a == 0) cm[t] = s_m[0] | s_m[1];
}

// sort key of the spatial order: (closest anchor, distance to it quantised to 1024 steps, point id)
__global__ void spatial_key_kernel(const PointMeta *__restrict__ meta, const float *__restrict__ D32, int64_t npad,
                                   int64_t n, float inv_scale, uint64_t *__restrict__ key)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        key[i] = ((uint64_t)row_bucket(meta[i], D32, npad, i, inv_scale) << 32) | (uint64_t)i;
}

__global__ void permute_D_kernel(const double *__restrict__ src, int64_t n, int na, const int32_t *__restrict__ order,
                                 double *__restrict__ dst)
{
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < (int64_t)na * n;
         q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = q / n, p = q % n;
        dst[q] = src[a * n + order[p]];
    }
}
}  // namespace annb

static int launch_scan_i32_i64(annb_ctx *c, const int32_t *in, int64_t *out, int64_t n, DevBuf &tmp)
{
    const int64_t nb = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    ANNB_TRY(tmp.ensure((size_t)(nb + 1) * 8));
    if (n == 0) {
        ANNB_CUDA(cudaMemsetAsync(out, 0, 8, c->stream));
        return ANNB_OK;
    }
    ANNB_LAUNCH(scan_block_sums_kernel, (int)nb, 1024, 0, c->stream, in, n, tmp.as<int64_t>());
    ANNB_LAUNCH(scan_sums_kernel, 1, 1024, 0, c->stream, tmp.as<int64_t>(), nb);
    ANNB_LAUNCH(scan_apply_kernel, (int)nb, 1024, 0, c->stream, in, n, tmp.as<int64_t>(), nb, out);
    return ANNB_OK;
}

// bring the per-tile entry lists up to date with the hash map (count, scan, fill: two streaming
// passes over the table)
static int tile_lists_rebuild(annb_index *ix)
{
    if (!ix->tl_dirty) return ANNB_OK;
    TraceScope _ts("  tile_lists_rebuild");
    annb_ctx *c = ix->ctx;
    const int64_t NT = ix->NT;
    ANNB_TRY(ix->tl_cnt.ensure((size_t)(NT + 1) * 4));
    ANNB_CUDA(cudaMemsetAsync(ix->tl_cnt.p, 0, (size_t)(NT + 1) * 4, c->stream));
    ANNB_LAUNCH(tl_count_kernel, grid_for_n(c, (int64_t)ix->hcap), 256, 0, c->stream, ix->htab.as<HashSlot>(),
                ix->hcap, ix->T, ix->tl_cnt.as<int32_t>());
    ANNB_TRY(launch_scan_i32_i64(c, ix->tl_cnt.as<int32_t>(), ix->tl_ptr.as<int64_t>(), NT, ix->scan_tmp));
    int64_t total = 0;
    ANNB_CUDA(cudaMemcpyAsync(&total, ix->tl_ptr.as<int64_t>() + NT, 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    ix->tl_entries = total;
    ANNB_TRY(ix->tl_code.ensure((size_t)(total + 64) * 4));
    ANNB_TRY(ix->tl_a.ensure((size_t)(total + 64) * 4));
    ANNB_TRY(ix->tl_b.ensure((size_t)(total + 64) * 4));
    ANNB_CUDA(cudaMemsetAsync(ix->tl_cnt.p, 0, (size_t)(NT + 1) * 4, c->stream));
    ANNB_LAUNCH(tl_fill_kernel, grid_for_n(c, (int64_t)ix->hcap), 256, 0, c->stream, ix->htab.as<HashSlot>(),
                ix->hcap, ix->T, ix->tl_ptr.as<long long>(), ix->tl_cnt.as<int32_t>(), ix->tl_code.as<uint32_t>(),
                ix->tl_a.as<float>(), ix->tl_b.as<float>());
    ix->tl_dirty = false;
    return ANNB_OK;
}

static int finish_anchors(annb_index *ix)
{
    annb_ctx *c = ix->ctx;
    ANNB_LAUNCH(build_meta_kernel, grid_for_n(c, ix->npad), 256, 0, c->stream, ix->D64.as<double>(),
                ix->n, ix->npad, ix->na, ix->P.locality, ix->P.loc_thresh, ix->D32.as<float>(),
                ix->Dpm.as<float>(), ix->dpitch, ix->meta.as<PointMeta>());
    if (!ix->A_host.empty())
        ANNB_LAUNCH(set_slots_kernel, 1, 32, 0, c->stream, ix->A_dev.as<int32_t>(),
                    (int)ix->A_host.size(), ix->meta.as<PointMeta>());
    ANNB_LAUNCH(tile_bounds_kernel, ix->T, 64, 0, c->stream, ix->D32.as<float>(), ix->meta.as<PointMeta>(), ix->n,
                ix->npad, ix->na, ix->tb_lo.as<float>(), ix->tb_hi.as<float>(), ix->tb_cm.as<uint64_t>());
    ix->have_anchors = true;
    ix->have_locality = false;
    return ANNB_OK;
}

// Spatial order: points sorted by (closest anchor, distance to it, id).  Points that are close to each
// other get close positions, so a 128-point tile covers a small region of the anchor-distance space and
// the sweeps can skip most tile pairs from the per-tile bounds.  order[new] = old.
ANNB_API int annb_index_spatial_order(annb_index *ix, int64_t *order)
{
    TraceScope _ts("annb_index_spatial_order");
    ANNB_REQUIRE(ix && order, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(ix->have_anchors, ANNB_ESTATE, "anchors not computed yet");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    const int64_t n = ix->n;
    ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
    ANNB_LAUNCH(max_f32_kernel, grid_for_n(c, (int64_t)ix->na * ix->npad), 256, 0, c->stream, ix->D32.as<float>(),
                (int64_t)ix->na * ix->npad, ix->counters.as<unsigned int>());
    float dmax = 0.0f;
    ANNB_CUDA(cudaMemcpyAsync(&dmax, ix->counters.p, 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    const float inv_scale = dmax > 0.0f && std::isfinite(dmax) ? 1023.0f / dmax : 0.0f;
    ANNB_TRY(ix->t0.ensure((size_t)n * 8));
    ANNB_LAUNCH(spatial_key_kernel, grid_for_n(c, n), 256, 0, c->stream, ix->meta.as<PointMeta>(), ix->D32.as<float>(),
                ix->npad, n, inv_scale, ix->t0.as<uint64_t>());
    std::vector<uint64_t> key(n);
    ANNB_CUDA(cudaMemcpyAsync(key.data(), ix->t0.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    // counting sort over the 64 x 1024 buckets; ids are visited in ascending order, so ties inside a
    // bucket stay ordered by point id (deterministic, identical on every rank)
    std::vector<int64_t> start((size_t)ROW_BUCKETS + 1, 0);
    for (int64_t i = 0; i < n; ++i) start[(size_t)(key[i] >> 32) + 1] += 1;
    for (int b = 0; b < ROW_BUCKETS; ++b) start[b + 1] += start[b];
    for (int64_t i = 0; i < n; ++i) order[start[(size_t)(key[i] >> 32)]++] = i;
    return ANNB_OK;
}

// Take over the anchors of `src` (an index over the same items in their original order) into `ix`, an
// index over the items gathered by `order` (annb_dataset_gather): D is permuted, A is renumbered.
ANNB_API int annb_index_adopt_anchors(annb_index *ix, annb_index *src, const int64_t *order)
{
    TraceScope _ts("annb_index_adopt_anchors");
    ANNB_REQUIRE(ix && src && order, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(src->have_anchors && src->n == ix->n && src->na == ix->na, ANNB_ESTATE,
                 "source index has no anchors or a different shape");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    const int64_t n = ix->n;
    std::vector<int32_t> o32(n), inv(n);
    for (int64_t p = 0; p < n; ++p) {
        ANNB_REQUIRE(order[p] >= 0 && order[p] < n, ANNB_EINVAL, "order is not a permutation");
        o32[p] = (int32_t)order[p];
        inv[order[p]] = (int32_t)p;
    }
    ANNB_TRY(ix->order_dev.ensure((size_t)n * 4));
    ANNB_CUDA(cudaMemcpyAsync(ix->order_dev.p, o32.data(), (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    ANNB_LAUNCH(permute_D_kernel, grid_for_n(c, (int64_t)ix->na * n), 256, 0, c->stream, src->D64.as<double>(), n, ix->na,
                ix->order_dev.as<int32_t>(), ix->D64.as<double>());
    ix->A_host.clear();
    for (int32_t a : src->A_host) ix->A_host.push_back(inv[a]);
    if (!ix->A_host.empty())
        ANNB_CUDA(cudaMemcpyAsync(ix->A_dev.p, ix->A_host.data(), ix->A_host.size() * 4, cudaMemcpyHostToDevice, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    ix->ordered = true;
    return finish_anchors(ix);
}

ANNB_API int annb_index_create(annb_ctx *c, const annb_dataset *ds, int metric,
                               const annb_index_params *P, annb_index **out)
{
    TraceScope _ts("annb_index_create");
    ANNB_REQUIRE(c && ds && P && out, ANNB_EINVAL, "NULL argument");
    ANNB_TRY(check_metric(ds, metric));
    ANNB_REQUIRE(P->n_anchors > 0 && P->n_anchors <= kMaxAnchors, ANNB_ERANGE,
                 "n_anchors=%d outside [1,%d]", P->n_anchors, kMaxAnchors);
    ANNB_REQUIRE(P->n_neighbors >= 2 && P->n_neighbors + 1 <= MAX_LIST, ANNB_ERANGE,
                 "n_neighbors=%d outside [2,%d]", P->n_neighbors, MAX_LIST - 1);
    ANNB_REQUIRE(P->locality >= 1 && P->loc_thresh >= 0 && P->loc_thresh <= 7, ANNB_ERANGE,
                 "locality=%d / loc_thresh=%d unsupported (loc_thresh must be <= 7)", P->locality,
                 P->loc_thresh);
    ANNB_REQUIRE(P->world >= 1 && P->rank >= 0 && P->rank < P->world, ANNB_EINVAL, "bad rank/world");
    ANNB_CUDA(cudaSetDevice(c->device));
    annb_index *ix = new annb_index();
    ix->ctx = c;
    ix->ds = ds;
    ix->metric = metric;
    ix->P = *P;
    ix->n = ds->n;
    ix->na = P->n_anchors;
    ix->T = (int)((ds->n + TILE - 1) / TILE);
    ix->npad = (int64_t)ix->T * TILE;
    ix->NT = (int64_t)ix->T * (ix->T + 1) / 2;
    int rc = ANNB_OK;
    do {
        if ((rc = ix->A_dev.ensure((size_t)ix->na * 4))) break;
        if ((rc = ix->D64.ensure((size_t)ix->na * ix->n * 8))) break;
        if ((rc = ix->D32.ensure((size_t)ix->na * ix->npad * 4))) break;
        ix->dpitch = (ix->na + 31) / 32 * 32;
        if ((rc = ix->Dpm.ensure((size_t)ix->npad * ix->dpitch * 4))) break;
        if ((rc = ix->meta.ensure((size_t)ix->npad * sizeof(PointMeta)))) break;
        if ((rc = ix->tl_ptr.ensure((size_t)(ix->NT + 1) * 8))) break;
        if ((rc = ix->tl_code.ensure(256))) break;
        if ((rc = ix->tb_lo.ensure((size_t)ix->T * kMaxAnchors * 4))) break;
        if ((rc = ix->tb_hi.ensure((size_t)ix->T * kMaxAnchors * 4))) break;
        if ((rc = ix->tb_cm.ensure((size_t)ix->T * 8))) break;
        if ((rc = ix->tl_a.ensure(256))) break;
        if ((rc = ix->tl_b.ensure(256))) break;
        if ((rc = ix->thresh.ensure((size_t)ix->npad * 4))) break;
        if ((rc = ix->counters.ensure(256))) break;
        if ((rc = ix->tiehist.ensure(65536 * 4))) break;
    } while (0);
    if (rc) {
        annb_index_destroy(ix);
        return rc;
    }
    ANNB_CUDA(cudaMemsetAsync(ix->tl_ptr.p, 0, (size_t)(ix->NT + 1) * 8, c->stream));  // every tile list empty
    ix->cull_enabled = getenv("ANNB_NO_CULL") == nullptr;  // test knob: sweeps without tile-level pruning
    ix->reduced_enabled = getenv("ANNB_NO_REDUCED") == nullptr;  // test knob: pruning without the reduced tile mode
    ix->scan_enabled = getenv("ANNB_NO_SCAN") == nullptr;        // test knob: pruning decided tile by tile, after the loads
    ANNB_TRY(hash_alloc(ix, 1 << 16));
    *out = ix;
    return ANNB_OK;
}

// size the known-pair store once for the whole fit (p_work * N(N-1)/2 evaluations plus the
// look-ahead pairs that may be tightened): growing it later means a rehash of GBs
ANNB_API int annb_index_reserve_pairs(annb_index *ix, int64_t n_pairs)
{
    TraceScope _ts("annb_index_reserve_pairs");
    ANNB_REQUIRE(ix != nullptr && n_pairs >= 0, ANNB_EINVAL, "bad argument");
    ANNB_CUDA(cudaSetDevice(ix->ctx->device));
    const int64_t extra = n_pairs - ix->hcount_ub;
    return extra > 0 ? hash_reserve(ix, extra) : ANNB_OK;
}

ANNB_API int annb_index_destroy(annb_index *ix)
{
    TraceScope _ts("annb_index_destroy");
    if (!ix) return ANNB_OK;
    cudaSetDevice(ix->ctx->device);
    cudaStreamSynchronize(ix->ctx->stream);
    DevBuf *all[] = {&ix->A_dev, &ix->D64, &ix->D32, &ix->Dpm, &ix->meta, &ix->scratch, &ix->htab, &ix->tl_ptr, &ix->tl_cnt, &ix->tl_code, &ix->tl_a, &ix->tl_b, &ix->scan_tmp, &ix->tb_lo, &ix->tb_hi, &ix->tb_cm, &ix->tb_cutmax, &ix->order_dev,
                     &ix->errs_dev, &ix->rank_dev, &ix->thresh, &ix->l2val, &ix->l2id, &ix->hist,
                     &ix->counters, &ix->emit_key, &ix->emit_lvl, &ix->sel_i, &ix->sel_j, &ix->nxt_i,
                     &ix->nxt_j, &ix->tiehist, &ix->tiekeys, &ix->pool_key, &ix->pool_dad, &ix->t0,
                     &ix->t1, &ix->t2, &ix->t3, &ix->t4, &ix->t5, &ix->t6, &ix->kptr, &ix->kids,
                     &ix->kds, &ix->kdeg, &ix->gptr, &ix->gJ, &ix->gsrc, &ix->row_order, &ix->twork, &ix->theavy,
                     &ix->tcut1, &ix->tcut2, &ix->trec, &ix->tcnt, &ix->tgat, &ix->tgcnt, &ix->l1part};
    for (DevBuf *b : all) b->release();
    delete ix;
    return ANNB_OK;
}

ANNB_API int annb_index_maxmin(annb_index *ix, int64_t first, int64_t *A)
{
    TraceScope _ts("annb_index_maxmin");
    ANNB_REQUIRE(ix != nullptr, ANNB_EINVAL, "index is NULL");
    ANNB_REQUIRE(first >= 0 && first < ix->n, ANNB_EINVAL, "first anchor out of range");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    const int32_t f = (int32_t)first;
    ANNB_CUDA(cudaMemcpyAsync(ix->A_dev.p, &f, 4, cudaMemcpyHostToDevice, c->stream));
    ANNB_TRY(maxmin_device(c, ix->ds, ix->metric, ix->na, ix->A_dev.as<int32_t>(),
                           ix->D64.as<double>(), ix->scratch));
    ix->A_host.resize(ix->na);
    ANNB_CUDA(cudaMemcpyAsync(ix->A_host.data(), ix->A_dev.p, (size_t)ix->na * 4,
                              cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    if (A)
        for (int k = 0; k < ix->na; ++k) A[k] = ix->A_host[k];
    return finish_anchors(ix);
}

ANNB_API int annb_index_set_anchors(annb_index *ix, const int64_t *A, int64_t nA, const double *D)
{
    TraceScope _ts("annb_index_set_anchors");
    ANNB_REQUIRE(ix && D, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(nA == 0 || nA == ix->na, ANNB_EINVAL, "A must be empty or have n_anchors entries");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    // D arrives (n, na) row-major (the reference's D.T); store anchor-major
    std::vector<double> Dam((size_t)ix->na * ix->n);
    for (int64_t j = 0; j < ix->n; ++j)
        for (int a = 0; a < ix->na; ++a) Dam[(size_t)a * ix->n + j] = D[j * ix->na + a];
    ANNB_CUDA(cudaMemcpyAsync(ix->D64.p, Dam.data(), Dam.size() * 8, cudaMemcpyHostToDevice, c->stream));
    ix->A_host.clear();
    for (int64_t k = 0; k < nA; ++k) {
        ANNB_REQUIRE(A[k] >= 0 && A[k] < ix->n, ANNB_EINVAL, "anchor id out of range");
        ix->A_host.push_back((int32_t)A[k]);
    }
    if (nA)
        ANNB_CUDA(cudaMemcpyAsync(ix->A_dev.p, ix->A_host.data(), (size_t)nA * 4,
                                  cudaMemcpyHostToDevice, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    return finish_anchors(ix);
}

ANNB_API int annb_index_get_D(annb_index *ix, double *D)
{
    ANNB_REQUIRE(ix && D, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(ix->have_anchors, ANNB_ESTATE, "anchors not computed yet");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    ANNB_TRY(ix->t0.ensure((size_t)ix->na * ix->n * 8));
    ANNB_TRY(transpose_D(c, ix->D64.as<double>(), ix->n, ix->na, ix->t0.as<double>()));
    ANNB_CUDA(cudaMemcpyAsync(D, ix->t0.p, (size_t)ix->na * ix->n * 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    return ANNB_OK;
}

ANNB_API int annb_index_locality(annb_index *ix, int64_t *n_candidates, int64_t *n_relaxed)
{
    TraceScope _ts("annb_index_locality");
    ANNB_REQUIRE(ix != nullptr, ANNB_EINVAL, "index is NULL");
    ANNB_REQUIRE(ix->have_anchors, ANNB_ESTATE, "anchors not computed yet");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    const int64_t n = ix->n;
    const int lt = ix->P.loc_thresh;
    ANNB_TRY(ix->t0.ensure((size_t)n * 8 * 4));
    const int grid = (int)((n + 255) / 256);
    ANNB_LAUNCH(locality_count_kernel, grid, 256, 0, c->stream, ix->meta.as<PointMeta>(), n, lt, 1,
                ix->t0.as<int32_t>(), ix->P.rank, ix->P.world);
    if (ix->P.world > 1) {  // row blocks are split over the ranks
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
        ANNB_TRY(ix->reduce(ix->t0.p, n * 8, ANNB_RED_I32 | ANNB_RED_DEVICE));
    }
    std::vector<int32_t> h((size_t)n * 8);
    ANNB_CUDA(cudaMemcpyAsync(h.data(), ix->t0.p, h.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    // t_i = min(loc_thresh, (loc_min+1)-th largest shared count)   (utils.py:472-480)
    const int64_t want = std::min<int64_t>(ix->P.loc_min, n - 1) + 1;
    std::vector<int8_t> t(n);
    int64_t relaxed = 0, total = 0;
    for (int64_t i = 0; i < n; ++i) {
        int64_t cum = 0;
        int ti = 0;
        for (int cval = lt; cval >= 0; --cval) {
            cum += h[i * 8 + cval];
            if (cum >= want) {
                ti = cval;
                break;
            }
        }
        t[i] = (int8_t)ti;
        if (ti < lt) ++relaxed;
    }
    std::vector<int32_t> &ncand = ix->ncand_host;
    ncand.assign(n, 0);
    if (relaxed) {
        ANNB_TRY(ix->t1.ensure((size_t)n));
        ANNB_CUDA(cudaMemcpyAsync(ix->t1.p, t.data(), (size_t)n, cudaMemcpyHostToDevice, c->stream));
        ANNB_LAUNCH(set_loc_t_kernel, grid_for_n(c, n), 256, 0, c->stream, ix->meta.as<PointMeta>(),
                    ix->t1.as<int8_t>(), n);
        ANNB_LAUNCH(locality_count_kernel, grid, 256, 0, c->stream, ix->meta.as<PointMeta>(), n, lt, 2,
                    ix->t0.as<int32_t>(), ix->P.rank, ix->P.world);
        if (ix->P.world > 1) {
            ANNB_CUDA(cudaStreamSynchronize(c->stream));
            ANNB_TRY(ix->reduce(ix->t0.p, n, ANNB_RED_I32 | ANNB_RED_DEVICE));
        }
        ANNB_CUDA(cudaMemcpyAsync(ncand.data(), ix->t0.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        for (int64_t i = 0; i < n; ++i) ncand[i] = h[i * 8 + lt] - 1;  // minus self
    }
    int64_t minc = INT64_MAX;
    for (int64_t i = 0; i < n; ++i) {
        total += ncand[i];
        minc = std::min<int64_t>(minc, ncand[i]);
    }
    // candidate pairs that touch an anchor: computed from the start (annchor.py:286-301)
    ix->anc_cand_host.assign(n, 0);
    ix->n_anchor_pairs = 0;
    if (!ix->A_host.empty()) {
        ANNB_LAUNCH(anchor_cand_kernel, grid_for_n(c, n), 256, 0, c->stream, ix->meta.as<PointMeta>(), n,
                    ix->A_dev.as<int32_t>(), (int)ix->A_host.size(), ix->t0.as<int32_t>());
        ANNB_CUDA(cudaMemcpyAsync(ix->anc_cand_host.data(), ix->t0.p, (size_t)n * 4, cudaMemcpyDeviceToHost,
                                  c->stream));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<char> is_anchor(n, 0);
        for (size_t k = 0; k < ix->A_host.size(); ++k) is_anchor[ix->A_host[k]] = 1;
        int64_t mixed = 0, both = 0;
        for (int64_t i = 0; i < n; ++i) {
            if (is_anchor[i]) {
                mixed += ncand[i] - ix->anc_cand_host[i];  // (anchor, non-anchor) pairs
                both += ix->anc_cand_host[i];              // (anchor, anchor) pairs, counted twice
            }
        }
        ix->n_anchor_pairs = mixed + both / 2;
    }
    ix->n_candidates = total / 2;
    ix->have_locality = true;
    if (n_candidates) *n_candidates = total / 2;
    if (n_relaxed) *n_relaxed = relaxed;
    // check_locality_size (utils.py:592-597, annchor.py:252-256)
    ANNB_REQUIRE(minc >= ix->P.n_neighbors, ANNB_ESTATE,
                 "Error: Not enough candidates in pool for all indices. Try again with higher locality.");
    return ANNB_OK;
}

static int upload_pairs(annb_index *ix, const int64_t *ij, const double *d, int64_t m)
{
    annb_ctx *c = ix->ctx;
    for (int64_t p = 0; p < 2 * m; ++p)
        ANNB_REQUIRE(ij[p] >= 0 && ij[p] < ix->n, ANNB_EINVAL, "pair index out of range");
    ANNB_TRY(ix->t0.ensure((size_t)m * 16));
    ANNB_TRY(ix->t1.ensure((size_t)m * 4));
    ANNB_TRY(ix->t2.ensure((size_t)m * 4));
    ANNB_TRY(ix->t3.ensure((size_t)m * 4));
    ANNB_CUDA(cudaMemcpyAsync(ix->t0.p, ij, (size_t)m * 16, cudaMemcpyHostToDevice, c->stream));
    double *dd = nullptr;
    if (d) {
        ANNB_TRY(ix->t4.ensure((size_t)m * 8));
        ANNB_CUDA(cudaMemcpyAsync(ix->t4.p, d, (size_t)m * 8, cudaMemcpyHostToDevice, c->stream));
        dd = ix->t4.as<double>();
    }
    ANNB_LAUNCH(split_ij32_kernel, grid_for_n(c, m), 256, 0, c->stream, ix->t0.as<int64_t>(), dd, m,
                ix->t1.as<int32_t>(), ix->t2.as<int32_t>(), ix->t3.as<float>());
    return ANNB_OK;
}

ANNB_API int annb_index_add_known(annb_index *ix, const int64_t *ij, const double *d, int64_t m)
{
    TraceScope _ts("annb_index_add_known");
    ANNB_REQUIRE(ix && (m == 0 || (ij && d)), ANNB_EINVAL, "NULL argument");
    if (m == 0) return ANNB_OK;
    ANNB_CUDA(cudaSetDevice(ix->ctx->device));
    ANNB_TRY(upload_pairs(ix, ij, d, m));
    ANNB_TRY(hash_insert(ix, ix->t1.as<int32_t>(), ix->t2.as<int32_t>(), ix->t3.as<float>(), nullptr,
                         KIND_KNOWN, m));
    ix->n_known += m;
    return ANNB_OK;
}

ANNB_API int annb_index_eval_pairs(annb_index *ix, const int64_t *ij, int64_t m, double *d)
{
    TraceScope _ts("annb_index_eval_pairs");
    ANNB_REQUIRE(ix && (m == 0 || ij), ANNB_EINVAL, "NULL argument");
    if (m == 0) return ANNB_OK;
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    ANNB_TRY(upload_pairs(ix, ij, nullptr, m));
    ANNB_TRY(pair_dists_f32_perm(c, ix->ds, ix->metric, ix->t1.as<int32_t>(), ix->t2.as<int32_t>(),
                                 nullptr, m, ix->t3.as<float>()));
    ANNB_TRY(hash_insert(ix, ix->t1.as<int32_t>(), ix->t2.as<int32_t>(), ix->t3.as<float>(), nullptr,
                         KIND_KNOWN, m));
    ix->n_known += m;
    if (d) {
        ANNB_TRY(ix->t4.ensure((size_t)m * 8));
        ANNB_TRY(pair_dists_f64(c, ix->ds, ix->metric, ix->t1.as<int32_t>(), ix->t2.as<int32_t>(), m,
                                ix->t4.as<double>()));
        ANNB_CUDA(cudaMemcpyAsync(d, ix->t4.p, (size_t)m * 8, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return ANNB_OK;
}

ANNB_API int annb_index_pair_features(annb_index *ix, const int64_t *ij, int64_t m, double *feat)
{
    TraceScope _ts("annb_index_pair_features");
    ANNB_REQUIRE(ix && (m == 0 || (ij && feat)), ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(ix->have_anchors, ANNB_ESTATE, "anchors not computed yet");
    if (m == 0) return ANNB_OK;
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    ANNB_TRY(upload_pairs(ix, ij, nullptr, m));
    ANNB_TRY(ix->t4.ensure((size_t)m * 12));
    ANNB_LAUNCH(pair_features_kernel, grid_for_n(c, m), 256, 0, c->stream, ix->view(),
                ix->t1.as<int32_t>(), ix->t2.as<int32_t>(), m, ix->t4.as<float>());
    std::vector<float> f((size_t)m * 3);
    ANNB_CUDA(cudaMemcpyAsync(f.data(), ix->t4.p, f.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t k = 0; k < f.size(); ++k) feat[k] = f[k];
    return ANNB_OK;
}

ANNB_API int annb_index_pair_state(annb_index *ix, const int64_t *ij, int64_t m, int32_t *kind, double *a,
                                   double *b)
{
    TraceScope _ts("annb_index_pair_state");
    ANNB_REQUIRE(ix && (m == 0 || (ij && kind)), ANNB_EINVAL, "NULL argument");
    if (m == 0) return ANNB_OK;
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    ANNB_TRY(upload_pairs(ix, ij, nullptr, m));
    ANNB_TRY(ix->t4.ensure((size_t)m * 8));
    ANNB_TRY(ix->t5.ensure((size_t)m * 4));
    ANNB_LAUNCH(pair_state_kernel, grid_for_n(c, m), 256, 0, c->stream, ix->view(), ix->t1.as<int32_t>(),
                ix->t2.as<int32_t>(), m, ix->t5.as<int32_t>(), ix->t4.as<float>());
    std::vector<float> f((size_t)m * 2);
    ANNB_CUDA(cudaMemcpyAsync(f.data(), ix->t4.p, f.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(kind, ix->t5.p, (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t k = 0; k < m; ++k) {
        if (a) a[k] = f[2 * k];
        if (b) b[k] = f[2 * k + 1];
    }
    return ANNB_OK;
}

ANNB_API int annb_index_set_model(annb_index *ix, const double *bins, const double *coef,
                                  const double *icpt, int64_t nb, const double *errs,
                                  const int64_t *eptr)
{
    TraceScope _ts("annb_index_set_model");
    ANNB_REQUIRE(ix && bins && coef && icpt, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(nb >= 1 && nb <= MAX_BINS, ANNB_ERANGE, "n_partitions=%lld outside [1,%d]",
                 (long long)nb, MAX_BINS);
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    Model &M = ix->model;
    memset(&M, 0, sizeof(M));
    M.nb = (int)nb;
    for (int b = 0; b <= nb; ++b) M.edge[b] = (float)bins[b];
    for (int b = 0; b < MAX_BINS; ++b) M.e2[b] = (b >= 1 && b < nb) ? 2.0f * M.edge[b] : INFINITY;
    for (int b = 0; b < nb; ++b) {
        M.c0[b] = (float)coef[3 * b];
        M.c1[b] = (float)coef[3 * b + 1];
        M.c2[b] = (float)coef[3 * b + 2];
        M.ic[b] = (float)icpt[b];
    }
    ix->errs_host.clear();
    ix->rank_host.clear();
    ix->nlevels = 0;
    if (errs && eptr) {
        const int64_t ne = eptr[nb];
        ANNB_REQUIRE(ne < (1 << 22), ANNB_ERANGE, "error tables too large (%lld)", (long long)ne);
        for (int b = 0; b <= nb; ++b) M.eoff[b] = (int)eptr[b];
        ix->errs_host.resize(ne);
        for (int64_t k = 0; k < ne; ++k) ix->errs_host[k] = (float)errs[k];
        // distinct probability values r/len over all labels, ascending -> level ids
        std::vector<double> vals;
        for (int b = 0; b < nb; ++b) {
            const int64_t len = eptr[b + 1] - eptr[b];
            ANNB_REQUIRE(len > 0, ANNB_EINVAL, "error table of label %d is empty", b);
            for (int64_t r = 0; r <= len; ++r) vals.push_back((double)r / (double)len);
        }
        std::vector<double> uniq = vals;
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        ANNB_REQUIRE(uniq.size() < 40000, ANNB_ERANGE,
                     "too many probability levels (%zu) for the shared-memory histogram", uniq.size());
        ix->nlevels = (int)uniq.size();
        ix->rank_host.resize(vals.size());
        for (size_t k = 0; k < vals.size(); ++k)
            ix->rank_host[k] = (uint16_t)(std::lower_bound(uniq.begin(), uniq.end(), vals[k]) - uniq.begin());
        ANNB_TRY(ix->errs_dev.ensure(ix->errs_host.size() * 4));
        ANNB_TRY(ix->rank_dev.ensure(ix->rank_host.size() * 2));
        ANNB_CUDA(cudaMemcpyAsync(ix->errs_dev.p, ix->errs_host.data(), ix->errs_host.size() * 4,
                                  cudaMemcpyHostToDevice, c->stream));
        ANNB_CUDA(cudaMemcpyAsync(ix->rank_dev.p, ix->rank_host.data(), ix->rank_host.size() * 2,
                                  cudaMemcpyHostToDevice, c->stream));
        ANNB_TRY(ix->hist.ensure((size_t)ix->nlevels * 4));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
    }
    ix->have_model = true;
    ix->have_thresh = false;
    return ANNB_OK;
}

// the row sweep (both triangles): all rows (split over the ranks, results summed), a column subset
// of all rows (pre-pass of the two-stage scheme), or an explicit list of row blocks on every rank
static int run_thresh_rows(annb_index *ix, int k2, int col_stride, float *cut2, const int32_t *rb_list, int n_rb)
{
    annb_ctx *c = ix->ctx;
    ANNB_TRY(tile_lists_rebuild(ix));
    ThreshArgs A;
    A.V = ix->view();
    A.M = ix->model;
    A.k1 = ix->P.n_neighbors + 1;
    A.k2 = k2;
    A.thresh = ix->thresh.as<float>();
    A.l2val = k2 > 0 ? ix->l2val.as<float>() : nullptr;
    A.l2id = k2 > 0 ? ix->l2id.as<int32_t>() : nullptr;
    A.col_stride = col_stride;
    A.col_phase = 0;
    A.band = (ix->ordered && col_stride > 1) ? 8 : 0;  // spatial order: a row's near neighbours sit around the diagonal
    A.cut2 = cut2;
    A.rb_list = rb_list;
    A.n_rb = n_rb;
    const bool split = ix->P.world > 1 && rb_list == nullptr;
    A.rank = split ? ix->P.rank : 0;
    A.world = split ? ix->P.world : 1;
    if (split) {  // rows of other ranks stay 0 so that a sum all-reduce assembles the result
        ANNB_CUDA(cudaMemsetAsync(ix->thresh.p, 0, (size_t)ix->npad * 4, c->stream));
        if (cut2) ANNB_CUDA(cudaMemsetAsync(cut2, 0, (size_t)ix->npad * 4, c->stream));
        if (k2 > 0 && col_stride == 1) {
            ANNB_CUDA(cudaMemsetAsync(ix->l2val.p, 0, (size_t)ix->n * k2 * 4, c->stream));
            ANNB_CUDA(cudaMemsetAsync(ix->l2id.p, 0, (size_t)ix->n * k2 * 4, c->stream));
        }
    }
    ANNB_TRY(launch_thresh_sweep(c, A));
    const int64_t rows = rb_list ? (int64_t)n_rb * TILE : ix->n / A.world;
    ix->pairs_swept += rows * (ix->n / col_stride);
    ix->sweeps += 1;
    if (split) {
        // every row block was computed by exactly one rank (the others hold zeros): a sum
        // all-reduce of the device buffers assembles the full vectors on every rank
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
        ANNB_TRY(ix->reduce(ix->thresh.p, ix->npad, ANNB_RED_F32 | ANNB_RED_DEVICE));
        if (cut2) ANNB_TRY(ix->reduce(cut2, ix->npad, ANNB_RED_F32 | ANNB_RED_DEVICE));
        if (k2 > 0 && col_stride == 1) {
            ANNB_TRY(ix->reduce(ix->l2val.p, ix->n * k2, ANNB_RED_F32 | ANNB_RED_DEVICE));
            ANNB_TRY(ix->reduce(ix->l2id.p, ix->n * k2, ANNB_RED_I32 | ANNB_RED_DEVICE));
        }
    }
    return ANNB_OK;
}

// thresh (annchor.py:399-404) and, with k2 > 0, the guarantee_nmin lists.  Large metric problems
// use the two-stage scheme (sweep_thresh.cu): column-subset pre-pass -> upper bounds, one visit per
// pair appending records to both endpoints, per-row selection; everything else the row sweep.
static int run_thresh(annb_index *ix, int k2)
{
    annb_ctx *c = ix->ctx;
    const int64_t n = ix->n;
    const int k1 = ix->P.n_neighbors + 1;
    if (k2 > 0) {
        ANNB_TRY(ix->l2val.ensure((size_t)n * k2 * 4));
        ANNB_TRY(ix->l2id.ensure((size_t)n * k2 * 4));
    }
    ANNB_CUDA(cudaEventRecord(c->ev0, c->stream));
    const bool force_rows = getenv("ANNB_THRESH_ROWS") != nullptr;  // test knob: row sweep only
    static const int S_env = getenv("ANNB_THRESH_STRIDE") ? atoi(getenv("ANNB_THRESH_STRIDE")) : 0;
    const int S = S_env > 0 ? S_env : 8;  // pre-pass visits every S-th column tile
    const bool two_stage = ix->P.is_metric && ix->T >= 16 * S && n < (1 << 30) && !force_rows;
    if (!two_stage) {
        ANNB_TRY(run_thresh_rows(ix, k2, 1, nullptr, nullptr, 0));
    } else {
        // 1. upper bounds from a column subset
        ANNB_TRY(ix->tcut1.ensure((size_t)ix->npad * 4));
        float *cut2 = nullptr;
        if (k2 > 0) {
            ANNB_TRY(ix->tcut2.ensure((size_t)ix->npad * 4));
            cut2 = ix->tcut2.as<float>();
        }
        ANNB_TRY(run_thresh_rows(ix, k2, S, cut2, nullptr, 0));
        ANNB_CUDA(cudaMemcpyAsync(ix->tcut1.p, ix->thresh.p, (size_t)ix->npad * 4, cudaMemcpyDeviceToDevice,
                                  c->stream));
        // 2. one visit per pair, records to both endpoints
        // record slots per point: ~S*(k1+k2) expected (the k-th smallest of 1/S of the columns is about
        // the S*k-th smallest overall), x2 (x4 without the second list) for the spread
        static const int R_env = getenv("ANNB_THRESH_RECORDS") ? atoi(getenv("ANNB_THRESH_RECORDS")) : 0;
        int R = R_env > 0 ? R_env : ((k2 > 0 ? 2 : 4) * S * (k1 + k2) + 31) / 32 * 32;
        // spatially ordered points: the typical row needs far fewer records (its neighbours are in the band
        // of the pre-pass), but the 0.1 % of rows whose blob is spread over several tiles need more
        if (ix->ordered && R_env <= 0) R = std::max(R, 1024);
        ANNB_TRY(ix->trec.ensure((size_t)n * R * 8));
        ANNB_TRY(ix->tcnt.ensure((size_t)n * 4));
        ANNB_CUDA(cudaMemsetAsync(ix->tcnt.p, 0, (size_t)n * 4, c->stream));
        ANNB_TRY(tile_lists_rebuild(ix));
        ThreshPairArgs P;
        P.V = ix->view();
        P.M = ix->model;
        P.cut1 = ix->tcut1.as<float>();
        P.cut2 = cut2;
        P.rec = ix->trec.as<uint2>();
        P.cnt = ix->tcnt.as<int32_t>();
        P.R = R;
        P.rank = ix->P.rank;
        P.world = ix->P.world;
        P.counters = nullptr;
        P.reduced = (P.V.cull && ix->reduced_enabled) ? 1 : 0;
        P.tcmax = nullptr;
        if (P.V.cull && ix->scan_enabled) {
            ANNB_TRY(ix->tb_cutmax.ensure((size_t)P.V.T * 4));
            ANNB_TRY(launch_tile_cutmax(c, P.cut1, P.cut2, P.V.T, ix->tb_cutmax.as<float>()));
            P.tcmax = ix->tb_cutmax.as<float>();
        }
        if (const char *dump = getenv("ANNB_DUMP_TILES")) {
            // debug: what the tile test of this sweep sees (analysed offline, tools/tile_prune_replay.py)
            static int dump_no = 0;
            char path[512];
            snprintf(path, sizeof path, "%s.%d.bin", dump, dump_no++);
            ANNB_CUDA(cudaStreamSynchronize(c->stream));
            const View V = ix->view();
            const int64_t T = V.T;
            std::vector<float> lo((size_t)T * kMaxAnchors), hi((size_t)T * kMaxAnchors), c1(ix->npad), c2(ix->npad, 0.0f);
            std::vector<uint64_t> cm(T);
            std::vector<long long> tl((size_t)ix->NT + 1);
            cudaMemcpy(lo.data(), V.tb_lo, lo.size() * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(hi.data(), V.tb_hi, hi.size() * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(cm.data(), V.tb_cm, cm.size() * 8, cudaMemcpyDeviceToHost);
            cudaMemcpy(c1.data(), P.cut1, c1.size() * 4, cudaMemcpyDeviceToHost);
            if (P.cut2) cudaMemcpy(c2.data(), P.cut2, c2.size() * 4, cudaMemcpyDeviceToHost);
            cudaMemcpy(tl.data(), V.tl_ptr, tl.size() * 8, cudaMemcpyDeviceToHost);
            if (FILE *f = fopen(path, "wb")) {
                const int64_t hdr[6] = {T, kMaxAnchors, V.na, ix->npad, P.cut2 ? 1 : 0, (int64_t)sizeof(Model)};
                fwrite(hdr, 8, 6, f);
                fwrite(&ix->model, sizeof(Model), 1, f);
                fwrite(lo.data(), 4, lo.size(), f);
                fwrite(hi.data(), 4, hi.size(), f);
                fwrite(cm.data(), 8, cm.size(), f);
                fwrite(c1.data(), 4, c1.size(), f);
                fwrite(c2.data(), 4, c2.size(), f);
                // store entries per tile as a byte (0 / 1) in upper-triangular order
                std::vector<uint8_t> he((size_t)ix->NT);
                for (int64_t t = 0; t < ix->NT; ++t) he[t] = tl[t + 1] > tl[t];
                fwrite(he.data(), 1, he.size(), f);
                fclose(f);
            }
        }
        if (g_trace) {
            ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
            P.counters = ix->counters.as<unsigned long long>();
        }
        ANNB_TRY(launch_thresh_pairs(c, P));
        if (g_trace) {
            unsigned long long tc[6];
            ANNB_CUDA(cudaMemcpyAsync(tc, ix->counters.p, 48, cudaMemcpyDeviceToHost, c->stream));
            ANNB_CUDA(cudaStreamSynchronize(c->stream));
            fprintf(stderr, "[annb-trace]   thresh pair sweep: tiles pruned %llu, reduced %llu (%llu with nothing to do, %llu rows), computed in full %llu (with store entries %llu) of %lld\n",
                    tc[0], tc[3], tc[4], tc[5], tc[2], tc[1], (long long)ix->NT);
        }
        ix->pairs_swept += n * (n - 1) / 2 / ix->P.world;
        ix->sweeps += 1;
        // 3. per-row selection (per-rank partial lists when the tiles are split)
        if (ix->P.world == 1) {
            ANNB_TRY(launch_thresh_select(c, P.rec, P.cnt, R, n, k1, k2, 1, ix->thresh.as<float>(), nullptr,
                                          k2 > 0 ? ix->l2val.as<float>() : nullptr,
                                          k2 > 0 ? ix->l2id.as<int32_t>() : nullptr));
        } else {
            // partial lists of all ranks side by side ([world][n][k] records, zero elsewhere, summed),
            // then the same selection over the world lists of each row
            const int W = ix->P.world, K = k1 + k2;
            ANNB_TRY(ix->tgat.ensure((size_t)W * n * K * 8));
            ANNB_TRY(ix->tgcnt.ensure((size_t)W * n * 4));
            ANNB_TRY(ix->t0.ensure((size_t)n * 4));
            ANNB_TRY(ix->t1.ensure((size_t)n * std::max(k2, 1) * 4));
            ANNB_TRY(ix->t2.ensure((size_t)n * std::max(k2, 1) * 4));
            ANNB_TRY(ix->l1part.ensure((size_t)n * k1 * 4));
            ANNB_TRY(launch_thresh_select(c, P.rec, P.cnt, R, n, k1, k2, 1, ix->t0.as<float>(), ix->l1part.as<float>(),
                                          ix->t1.as<float>(), ix->t2.as<int32_t>()));
            ANNB_CUDA(cudaMemsetAsync(ix->tgat.p, 0, (size_t)W * n * K * 8, c->stream));
            ANNB_CUDA(cudaMemsetAsync(ix->tgcnt.p, 0, (size_t)W * n * 4, c->stream));
            ANNB_LAUNCH(pack_partial_lists_kernel, grid_for_n(c, n * K), 256, 0, c->stream, ix->t0.as<float>(),
                        ix->t1.as<float>(), ix->t2.as<int32_t>(), ix->l1part.as<float>(), n, k1, k2,
                        ix->tgat.as<uint2>() + (size_t)ix->P.rank * n * K,
                        ix->tgcnt.as<int32_t>() + (size_t)ix->P.rank * n);
            ANNB_CUDA(cudaStreamSynchronize(c->stream));
            ANNB_TRY(ix->reduce(ix->tgat.p, (int64_t)W * n * K * 2, ANNB_RED_I32 | ANNB_RED_DEVICE));
            ANNB_TRY(ix->reduce(ix->tgcnt.p, (int64_t)W * n, ANNB_RED_I32 | ANNB_RED_DEVICE));
            ANNB_TRY(launch_thresh_select(c, ix->tgat.as<uint2>(), ix->tgcnt.as<int32_t>(), K, n, k1, k2, W,
                                          ix->thresh.as<float>(), nullptr, k2 > 0 ? ix->l2val.as<float>() : nullptr,
                                          k2 > 0 ? ix->l2id.as<int32_t>() : nullptr));
        }
        // 4. rows whose record list overflowed: recompute their row blocks with the row sweep
        std::vector<int32_t> cnt_h(n);
        ANNB_CUDA(cudaMemcpyAsync(cnt_h.data(), ix->tcnt.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
        std::vector<uint32_t> bad((size_t)ix->T, 0);
        for (int64_t i = 0; i < n; ++i)
            if (cnt_h[i] > R) bad[i / TILE] = 1;
        ANNB_TRY(ix->reduce(bad.data(), (int64_t)bad.size(), ANNB_RED_I32));  // same list on every rank
        std::vector<int32_t> rbs;
        for (int t = 0; t < ix->T; ++t)
            if (bad[t]) rbs.push_back(t);
        if (g_trace) {
            std::vector<int32_t> srt(cnt_h);
            std::sort(srt.begin(), srt.end());
            int64_t over = 0, over_anchor = 0;
            std::vector<char> is_a(n, 0);
            for (int32_t a : ix->A_host) is_a[a] = 1;
            for (int64_t i = 0; i < n; ++i)
                if (cnt_h[i] > R) {
                    ++over;
                    over_anchor += is_a[i];
                }
            fprintf(stderr, "[annb-trace]   two-stage thresh: R %d, %zu of %d row blocks overflowed; records per row p50 %d p99 %d "
                            "p99.9 %d max %d; %lld rows over R (%lld anchors)\n",
                    R, rbs.size(), ix->T, srt[n / 2], srt[n * 99 / 100], srt[n * 999 / 1000], srt[n - 1], (long long)over,
                    (long long)over_anchor);
        }
        if (!rbs.empty()) {
            ANNB_TRY(ix->t3.ensure(rbs.size() * 4));
            ANNB_CUDA(cudaMemcpyAsync(ix->t3.p, rbs.data(), rbs.size() * 4, cudaMemcpyHostToDevice, c->stream));
            ANNB_TRY(run_thresh_rows(ix, k2, 1, nullptr, ix->t3.as<int32_t>(), (int)rbs.size()));
        }
    }
    ANNB_CUDA(cudaEventRecord(c->ev1, c->stream));
    ANNB_CUDA(cudaEventSynchronize(c->ev1));
    ix->have_thresh = true;
    return ANNB_OK;
}

ANNB_API int annb_index_row_thresh(annb_index *ix, double *thresh)
{
    TraceScope _ts("annb_index_row_thresh");
    ANNB_REQUIRE(ix != nullptr, ANNB_EINVAL, "index is NULL");
    ANNB_REQUIRE(ix->have_anchors && ix->have_model, ANNB_ESTATE, "set anchors and model first");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    ANNB_TRY(run_thresh(ix, 0));
    if (thresh) {
        std::vector<float> h(ix->n);
        ANNB_CUDA(cudaMemcpyAsync(h.data(), ix->thresh.p, (size_t)ix->n * 4, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
        for (int64_t i = 0; i < ix->n; ++i) thresh[i] = h[i];
    }
    return ANNB_OK;
}

// guarantee_nmin (utils.py:606-621).  The thresh sweep runs with a second per-row list (the
// nmin+1 smallest not-computed predictions with ids); the reference's serial row loop, in which
// later rows see the -1 marks of earlier rows, then runs on the host over those short lists.
ANNB_API int annb_index_guarantee_nmin(annb_index *ix, int64_t nmin, int64_t *n_forced)
{
    TraceScope _ts("annb_index_guarantee_nmin");
    ANNB_REQUIRE(ix != nullptr, ANNB_EINVAL, "index is NULL");
    ANNB_REQUIRE(ix->have_anchors && ix->have_model && ix->have_locality, ANNB_ESTATE,
                 "anchors, locality and model must be set first");
    ANNB_REQUIRE(nmin >= 0 && nmin + 1 <= MAX_LIST, ANNB_ERANGE,
                 "nmin=%lld too large for the device row lists (max %d)", (long long)nmin, MAX_LIST - 1);
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    const int64_t n = ix->n;
    if (n_forced) *n_forced = 0;
    // computed pairs per row = anchor pairs among the candidates + exactly known pairs
    ANNB_TRY(ix->kdeg.ensure((size_t)(n + 1) * 4));
    ANNB_CUDA(cudaMemsetAsync(ix->kdeg.p, 0, (size_t)(n + 1) * 4, c->stream));
    ANNB_LAUNCH(known_degree_kernel, grid_for_n(c, (int64_t)ix->hcap), 256, 0, c->stream,
                ix->htab.as<HashSlot>(), ix->hcap, ix->kdeg.as<int32_t>());
    std::vector<int32_t> ncomp(n);
    ANNB_CUDA(cudaMemcpyAsync(ncomp.data(), ix->kdeg.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<char> is_anchor(n, 0);
    for (size_t k = 0; k < ix->A_host.size(); ++k) is_anchor[ix->A_host[k]] = 1;
    bool any = false;
    for (int64_t i = 0; i < n; ++i) {
        ncomp[i] += is_anchor[i] ? ix->ncand_host[i] : ix->anc_cand_host[i];
        any |= ncomp[i] < nmin;
    }
    if (!any) {  // every row already has nmin computed pairs: only thresh is needed
        return run_thresh(ix, 0);
    }
    const int k2 = (int)nmin + 1;
    ANNB_TRY(run_thresh(ix, k2));
    std::vector<float> lv((size_t)n * k2);
    std::vector<int32_t> li((size_t)n * k2);
    ANNB_CUDA(cudaMemcpyAsync(lv.data(), ix->l2val.p, lv.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(li.data(), ix->l2id.p, li.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    std::vector<int64_t> fij;
    std::unordered_map<int32_t, std::vector<int32_t>> marks;  // row -> earlier rows that forced (row', row)
    for (int64_t i = 0; i < n; ++i) {
        const int64_t n_todo = nmin - ncomp[i];
        if (n_todo <= 0) continue;
        const auto it = marks.find((int32_t)i);
        const std::vector<int32_t> *mk = it == marks.end() ? nullptr : &it->second;
        const int64_t m_i = mk ? (int64_t)mk->size() : 0;
        if (m_i >= n_todo + 1) continue;  // the (n_todo+1)-th smallest is already a -1 mark
        const float *v = &lv[(size_t)i * k2];
        const int32_t *id = &li[(size_t)i * k2];
        std::vector<int> un;  // unmarked entries of the row's list, ascending
        for (int k = 0; k < k2; ++k) {
            if (id[k] < 0) continue;
            bool marked = false;
            if (mk)
                for (int32_t r : *mk)
                    if (r == id[k]) {
                        marked = true;
                        break;
                    }
            if (!marked) un.push_back(k);
        }
        const int64_t need = n_todo + 1 - m_i;  // rank (1-based) of kth among the unmarked values
        if ((int64_t)un.size() < need) continue;
        const float kth = v[un[need - 1]];
        for (int64_t q = 0; q < need - 1; ++q) {
            if (!(v[un[q]] < kth)) break;  // strict <: ties with kth are dropped (utils.py:602-603)
            const int32_t j = id[un[q]];
            fij.push_back(std::min<int64_t>(i, j));
            fij.push_back(std::max<int64_t>(i, j));
            if (j > i) marks[j].push_back((int32_t)i);
        }
    }
    const int64_t m = (int64_t)fij.size() / 2;
    if (m) {
        ANNB_TRY(upload_pairs(ix, fij.data(), nullptr, m));
        ANNB_TRY(hash_insert(ix, ix->t1.as<int32_t>(), ix->t2.as<int32_t>(), nullptr, nullptr,
                             KIND_FORCED, m));
        ix->has_forced = true;
    }
    if (n_forced) *n_forced = m;
    return ANNB_OK;
}

// ---- selection --------------------------------------------------------------------------------
static void fill_efloor(const annb_index *ix, int floor_level, float *efloor, float *ef_min)
{
    const Model &M = ix->model;
    for (int b = 0; b < MAX_BINS; ++b) efloor[b] = INFINITY;
    float mn = INFINITY;
    for (int b = 0; b < M.nb; ++b) {
        const int len = M.eoff[b + 1] - M.eoff[b];
        int rmin = -1;
        for (int r = 0; r <= len; ++r)
            if (ix->rank_host[M.eoff[b] + b + r] >= floor_level) {
                rmin = r;
                break;
            }
        if (rmin < 0) efloor[b] = INFINITY;
        else if (rmin == 0) efloor[b] = -INFINITY;
        else efloor[b] = ix->errs_host[M.eoff[b] + rmin - 1];
        mn = std::min(mn, efloor[b]);
    }
    *ef_min = ix->P.is_metric ? mn : -INFINITY;  // d >= lower bound only holds for metrics
}

static int run_score(annb_index *ix, int floor_level, uint64_t floor_mix_thr, int stride, bool emit,
                     int64_t emit_cap, std::vector<uint64_t> &hist, unsigned long long cnt[8])
{
    annb_ctx *c = ix->ctx;
    ANNB_TRY(tile_lists_rebuild(ix));
    ScoreArgs A;
    A.V = ix->view();
    A.M = ix->model;
    A.thresh = ix->thresh.as<float>();
    A.errs = ix->errs_dev.as<float>();
    A.ranktab = ix->rank_dev.as<uint16_t>();
    A.nlevels = ix->nlevels;
    A.n_errs = (int)ix->errs_host.size();
    A.floor_level = floor_level;
    A.floor_mix_thr = floor_mix_thr;
    A.tie_salt = ix->tie_salt;
    fill_efloor(ix, floor_level, A.efloor, &A.ef_min);
    {
        float dummy;
        fill_efloor(ix, floor_level + 1, A.efloor_hi, &dummy);
    }
    A.wide_floor = floor_mix_thr != ~0ull ? 1 : 0;
    A.has_forced = ix->has_forced ? 1 : 0;
    A.hist = ix->hist.as<uint32_t>();
    A.counters = ix->counters.as<unsigned long long>();
    A.emit_key = nullptr;
    A.emit_lvl = nullptr;
    A.emit_cap = 0;
    if (emit) {
        ANNB_TRY(ix->emit_key.ensure((size_t)emit_cap * 8));
        ANNB_TRY(ix->emit_lvl.ensure((size_t)emit_cap * 2));
        A.emit_key = ix->emit_key.as<uint64_t>();
        A.emit_lvl = ix->emit_lvl.as<uint16_t>();
        A.emit_cap = (unsigned long long)emit_cap;
    }
    const int64_t nq = (ix->NT - ix->P.rank + ix->P.world - 1) / ix->P.world;
    A.q_begin = 0;
    A.q_end = nq;
    A.q_stride = stride;
    A.rank = ix->P.rank;
    A.world = ix->P.world;
    A.reduced = (A.V.cull && ix->reduced_enabled) ? 1 : 0;
    A.tcmax = nullptr;
    if (A.V.cull && ix->scan_enabled) {
        ANNB_TRY(ix->tb_cutmax.ensure((size_t)A.V.T * 4));
        ANNB_TRY(launch_tile_cutmax(c, A.thresh, nullptr, A.V.T, ix->tb_cutmax.as<float>()));
        A.tcmax = ix->tb_cutmax.as<float>();
    }
    ANNB_CUDA(cudaMemsetAsync(ix->hist.p, 0, (size_t)ix->nlevels * 4, c->stream));
    ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
    ANNB_CUDA(cudaEventRecord(c->ev0, c->stream));
    ANNB_TRY(launch_score_sweep(c, A));
    ANNB_CUDA(cudaEventRecord(c->ev1, c->stream));
    std::vector<uint32_t> h32(ix->nlevels);
    ANNB_CUDA(cudaMemcpyAsync(h32.data(), ix->hist.p, (size_t)ix->nlevels * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(cnt, ix->counters.p, 56, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    hist.assign(h32.begin(), h32.end());
    if (g_trace)
        fprintf(stderr, "[annb-trace]   score sweep: stride %d floor %d pairs-in-phase2 %llu (flagged %llu) not-computed %llu emitted %llu, tiles pruned %llu, reduced %llu, pairs computed %llu\n",
                stride, floor_level, cnt[3], cnt[4], cnt[1], cnt[0], cnt[5], cnt[6], cnt[2]);
    ANNB_TRY(ix->reduce(hist.data(), (int64_t)hist.size(), ANNB_RED_U64));  // global level counts
    float ms = 0;
    ANNB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    // pairs whose bounds / prediction were actually computed (tiles skipped by the tile-level test excluded)
    const int64_t pairs = (int64_t)cnt[2];
    if (stride == 1) {  // accumulated over the full scoring sweeps of this index (bench.py: roofline)
        ix->last_sweep_ms += ms;
        ix->last_sweep_pairs += pairs;
    }
    ix->pairs_swept += pairs;
    ix->sweeps += 1;
    return ANNB_OK;
}

// exact t-th smallest (1-based) mixed key among the emitted entries of one level
static int tie_threshold(annb_index *ix, int64_t E, int level, int64_t level_count, int64_t t, uint64_t *thr)
// level_count = number of EMITTED pairs at this level (all of them, or -- at the floor level -- those
// whose mixed key passed the emission threshold, which are exactly the smallest ones)
{
    annb_ctx *c = ix->ctx;
    ANNB_TRY(ix->tiekeys.ensure((size_t)(level_count + 32) * 8));
    ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
    ANNB_LAUNCH(compact_level_kernel, grid_for_n(c, E, 256 * PE_ITEMS), 256, 0, c->stream, ix->emit_key.as<uint64_t>(),
                ix->emit_lvl.as<uint16_t>(), E, ix->tie_salt, level, ix->tiekeys.as<uint64_t>(), level_count + 32,
                ix->counters.as<unsigned long long>());
    unsigned long long found = 0;
    ANNB_CUDA(cudaMemcpyAsync(&found, ix->counters.p, 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    const int64_t found_local = (int64_t)found;
    uint64_t found_all = found;
    ANNB_TRY(ix->reduce(&found_all, 1, ANNB_RED_U64));
    ANNB_REQUIRE((int64_t)found_all == level_count, ANNB_ESTATE,
                 "tie compaction found %llu emitted pairs at level %d, the level histogram says %lld (E=%lld)",
                 (unsigned long long)found_all, level, (long long)level_count, (long long)E);
    uint64_t prefix = 0;
    std::vector<uint32_t> h32(65536);
    std::vector<uint64_t> h(65536);
    for (int shift = 48; shift >= 0; shift -= 16) {
        ANNB_CUDA(cudaMemsetAsync(ix->tiehist.p, 0, 65536 * 4, c->stream));
        if (found_local > 0)
            ANNB_LAUNCH(digit_hist_kernel, grid_for_n(c, found_local), 256, 0, c->stream,
                        ix->tiekeys.as<uint64_t>(), found_local, prefix, shift, ix->tiehist.as<uint32_t>());
        ANNB_CUDA(cudaMemcpyAsync(h32.data(), ix->tiehist.p, 65536 * 4, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
        h.assign(h32.begin(), h32.end());
        ANNB_TRY(ix->reduce(h.data(), 65536, ANNB_RED_U64));
        int d = 0;
        for (; d < 65536; ++d) {
            if (t <= (int64_t)h[d]) break;
            t -= h[d];
        }
        ANNB_REQUIRE(d < 65536, ANNB_ESTATE, "tie selection ran past the histogram");
        prefix = (prefix << 16) | (uint64_t)d;
    }
    *thr = prefix;
    return ANNB_OK;
}

// largest level L >= lo_level with sum_{l >= L} h[l] >= target (lo_level if none)
static int level_cut(const std::vector<uint64_t> &h, int lo_level, int64_t target, int64_t *cum_at)
{
    int64_t cum = 0;
    for (int L = (int)h.size() - 1; L > lo_level; --L) {
        cum += h[L];
        if (cum >= target) {
            if (cum_at) *cum_at = cum;
            return L;
        }
    }
    cum += h[lo_level];
    if (cum_at) *cum_at = cum;
    return lo_level;
}

ANNB_API int annb_index_select(annb_index *ix, int64_t n_refine, int64_t lookahead,
                               int64_t *n_selected, int64_t *n_next)
{
    TraceScope _ts("annb_index_select");
    ANNB_REQUIRE(ix != nullptr, ANNB_EINVAL, "index is NULL");
    ANNB_REQUIRE(ix->have_thresh && ix->nlevels > 0 && ix->have_locality, ANNB_ESTATE,
                 "locality, set_model (with error tables) and row_thresh must run before select");
    ANNB_REQUIRE(n_refine >= 0 && lookahead >= 1, ANNB_EINVAL, "bad n_refine / lookahead");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    ix->n_sel = ix->n_next = 0;
    ix->n_selects += 1;
    static const uint64_t salt0 = getenv("ANNB_TIE_SALT") ? strtoull(getenv("ANNB_TIE_SALT"), nullptr, 0) : 0ull;
    ix->tie_salt = mix64(0x9E3779B97F4A7C15ull * (uint64_t)ix->n_selects + salt0);  // ANNB_TIE_SALT: test knob
    const int64_t nq = (ix->NT - ix->P.rank + ix->P.world - 1) / ix->P.world;
    const int64_t n_nc = ix->n_not_computed();  // exact: candidates - anchor pairs - known
    const int64_t want1 = n_refine, want2 = n_refine * lookahead;
    // annchor.py:444-457: n_refine >= len(prob) -> everything is candidate AND next;
    // n_refine*lookahead >= len(prob) -> large_part is everything
    int64_t sel_target = std::min<int64_t>(want1, n_nc);
    int64_t tot_target = want1 >= n_nc ? n_nc : std::min<int64_t>(want2, n_nc);

    std::vector<uint64_t> h1, h2;
    unsigned long long c1v[8] = {0, 0, 0, 0, 0, 0, 0, 0}, c2v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    // Level 0 is probability 0.  On small problems it takes part like any other level (the
    // reference's argpartition then picks arbitrary probability-0 pairs).  On large problems
    // emitting the probability-0 bulk is pointless and Theta(N^2): the cut never goes below level 1,
    // and if fewer positive-probability pairs exist than asked for, the look-ahead set (then the
    // selection) is truncated to what exists.
    const bool small = tot_target >= n_nc || n_nc <= (int64_t)64000000;
    const int min_floor = small ? 0 : 1;
    int floor_level = min_floor;
    // Probability levels are heavily tied (<= ~716 values per label), so the level that contains the
    // cut can hold many times more pairs than are needed from it.  Ties at a cut are resolved by the
    // smallest splitmix64(pair key); the sweep therefore emits, AT the floor level, only the pairs whose
    // mixed key is below `floor_thr` -- a superset of the ones the cut can take.
    uint64_t floor_thr = ~0ull;
    int64_t cap = 0;
    // floor level + emission threshold + list capacity from (estimated) level counts h * scale
    // (valid for levels >= lo_level only: an emitting pass does not count the levels below its floor)
    auto plan = [&](const std::vector<uint64_t> &h, int lo_level, double scale, double slack) {
        const double target = slack * (double)tot_target + (scale > 1.0 ? 65536.0 : 0.0);
        double above = 0;
        int L = ix->nlevels - 1;
        for (; L > lo_level; --L) {
            if (above + (double)h[L] * scale >= target) break;
            above += (double)h[L] * scale;
        }
        const double at = (double)h[L] * scale;
        double need = target - above;
        if (need < 0) need = 0;
        double want = need + 6.0 * sqrt(need) + 64.0;  // binomial slack of the hash threshold
        if (scale > 1.0) want = want * 1.1 + 4096.0;   // + pilot sampling error
        floor_level = L;
        if (want >= 0.999 * at) {
            floor_thr = ~0ull;
            want = at;
        } else {
            floor_thr = (uint64_t)((want / at) * 18446744073709551615.0);
        }
        cap = (int64_t)((above + want) * (scale > 1.0 ? 1.15 : 1.0)) + (int64_t)(8.0 * sqrt(above + want)) + 65536;
        if (cap > n_nc + 1024) cap = n_nc + 1024;
    };
    if (tot_target >= n_nc) {
        cap = n_nc + 1024;  // everything is taken: no pilot needed
    } else {
        // pass 1: level histogram -- exact for small problems, a pilot over every s-th tile otherwise
        const int stride = nq > 8192 ? (int)(nq / 2048) : 1;
        ANNB_TRY(run_score(ix, min_floor, ~0ull, stride, false, 0, h1, c1v));
        if (stride == 1) plan(h1, min_floor, 1.0, 1.0);
        else plan(h1, min_floor, (double)nq / (double)((nq + stride - 1) / stride), 1.25);
    }
    // pass 2: exact histogram of the levels >= floor, with emission; re-plan if the pilot was off
    int64_t E = 0, E_all = 0, at_floor_emitted = 0;
    for (int attempt = 0;; ++attempt) {
        ANNB_REQUIRE(attempt < 5, ANNB_ESTATE, "selection did not converge");
        // `cap` counts pairs over all ranks; a rank's own list gets its share plus slack
        const int64_t cap_local = ix->P.world > 1 ? cap / ix->P.world + cap / (4 * ix->P.world) + (1 << 16) : cap;
        ANNB_TRY(run_score(ix, floor_level, floor_thr, 1, true, cap_local, h2, c2v));
        E = (int64_t)c2v[0];
        uint64_t red[2] = {E > cap_local ? 1ull : 0ull, (uint64_t)E};
        ANNB_TRY(ix->reduce(red, 2, ANNB_RED_U64));  // every rank must take the same branch
        E_all = (int64_t)red[1];
        int64_t cum = 0;
        for (int L = floor_level; L < ix->nlevels; ++L) cum += h2[L];
        const int64_t above = cum - (int64_t)h2[floor_level];
        const bool wide = floor_thr != ~0ull;  // h2[floor] then counts only the pairs that passed the hash
        const double frac = (double)floor_thr / 18446744073709551615.0;
        if (red[0]) {
            if (g_trace) fprintf(stderr, "[annb-trace]   select attempt %d: overflow E %lld cap %lld floor %d\n", attempt,
                                 (long long)E_all, (long long)cap, floor_level);
            // more pairs emitted than planned: the histogram above the floor is exact, re-plan from it
            // (the floor level's population is extrapolated from the observed hash pass rate)
            std::vector<uint64_t> hc(h2);
            if (wide) hc[floor_level] = (uint64_t)((double)h2[floor_level] / std::max(frac, 1e-12));
            plan(hc, floor_level, wide ? 1.0000001 : 1.0, 1.0);
            continue;
        }
        at_floor_emitted = E_all - above;
        if (g_trace)
            fprintf(stderr, "[annb-trace]   select attempt %d: floor %d thr %.6f cap %lld E %lld cum %lld above %lld "
                            "at_floor %llu emitted_at_floor %lld targets %lld/%lld n_nc %lld\n",
                    attempt, floor_level, frac, (long long)cap, (long long)E_all, (long long)cum, (long long)above,
                    (unsigned long long)h2[floor_level], (long long)at_floor_emitted, (long long)sel_target,
                    (long long)tot_target, (long long)n_nc);
        if (wide && above < tot_target && tot_target - above > at_floor_emitted) {
            // the hash threshold at the floor level let too few pairs through: scale it by the
            // observed pass rate (+10 %), everything at the floor level on the last attempts
            const double need = (double)(tot_target - above);
            const double nf = frac * need / (double)std::max<int64_t>(at_floor_emitted, 1) * 1.1 +
                              256.0 * frac / (double)std::max<int64_t>(at_floor_emitted, 1);
            const double pass_est = need * 1.1 + 256.0;
            if (nf >= 0.999 || attempt >= 2) {
                floor_thr = ~0ull;
                cap = n_nc + 1024;  // population at the floor level unknown: worst case
                cap = std::min<int64_t>(cap, (int64_t)((double)above + (double)at_floor_emitted / std::max(frac, 1e-12) * 1.2) + (1 << 20));
            } else {
                floor_thr = (uint64_t)(nf * 18446744073709551615.0);
                cap = std::min<int64_t>((int64_t)((double)above + pass_est * 1.1) + 65536, n_nc + 1024);
            }
            continue;
        }
        if (floor_level > min_floor && cum < tot_target) {
            // floor too high: exact full histogram, exact floor
            ANNB_TRY(run_score(ix, min_floor, ~0ull, 1, false, 0, h1, c1v));
            plan(h1, min_floor, 1.0, 1.0);
            continue;
        }
        if (cum < tot_target) {  // large problem, not enough positive-probability pairs
            tot_target = cum;
            sel_target = std::min<int64_t>(sel_target, cum);
        }
        break;
    }
    // ---- cut the emitted list: level first, mixed pair key among ties (deterministic) ----
    // he[L] = emitted pairs at level L
    std::vector<uint64_t> he(h2);
    he[floor_level] = (uint64_t)std::max<int64_t>(at_floor_emitted, 0);
    int cl1 = ix->nlevels, cl2 = ix->nlevels;
    int64_t t1 = 0, t2 = 0;
    {
        int64_t above = 0;
        for (int L = ix->nlevels - 1; L >= floor_level; --L) {
            if (cl1 == ix->nlevels && sel_target > 0 && above + (int64_t)h2[L] >= sel_target) {
                cl1 = L;
                t1 = sel_target - above;
            }
            if (cl2 == ix->nlevels && tot_target > 0 && above + (int64_t)h2[L] >= tot_target) {
                cl2 = L;
                t2 = tot_target - above;
            }
            above += h2[L];
        }
        ANNB_REQUIRE(above >= tot_target, ANNB_ESTATE,
                     "scoring sweep found %lld pairs, %lld expected (n_not_computed=%lld)", (long long)above,
                     (long long)tot_target, (long long)n_nc);
    }
    uint64_t thr1 = ~0ull, thr2 = ~0ull;
    if (cl1 < ix->nlevels && t1 < (int64_t)he[cl1]) ANNB_TRY(tie_threshold(ix, E, cl1, he[cl1], t1, &thr1));
    if (cl2 < ix->nlevels && t2 < (int64_t)he[cl2]) ANNB_TRY(tie_threshold(ix, E, cl2, he[cl2], t2, &thr2));
    const int64_t cap_sel = std::min<int64_t>(sel_target, E) + 16;
    const int64_t cap_next = std::min<int64_t>(tot_target - sel_target, E) + 16;
    ANNB_TRY(ix->sel_i.ensure((size_t)cap_sel * 4));
    ANNB_TRY(ix->sel_j.ensure((size_t)cap_sel * 4));
    ANNB_TRY(ix->nxt_i.ensure((size_t)cap_next * 4));
    ANNB_TRY(ix->nxt_j.ensure((size_t)cap_next * 4));
    ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
    if (E > 0)
        ANNB_LAUNCH(partition_emitted_kernel, grid_for_n(c, E, 256 * PE_ITEMS), 256, 0, c->stream,
                    ix->emit_key.as<uint64_t>(), ix->emit_lvl.as<uint16_t>(), E, ix->tie_salt, cl1, thr1, cl2, thr2,
                    ix->sel_i.as<int32_t>(), ix->sel_j.as<int32_t>(), ix->nxt_i.as<int32_t>(),
                    ix->nxt_j.as<int32_t>(), ix->counters.as<unsigned long long>(), cap_sel, cap_next);
    unsigned long long pc[2] = {0, 0};
    ANNB_CUDA(cudaMemcpyAsync(pc, ix->counters.p, 16, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    uint64_t pg[2] = {pc[0], pc[1]};
    ANNB_TRY(ix->reduce(pg, 2, ANNB_RED_U64));
    ANNB_REQUIRE((int64_t)pg[0] == sel_target && (int64_t)pg[1] == tot_target - sel_target, ANNB_ESTATE,
                 "selection cut produced %llu/%llu pairs for targets %lld/%lld",
                 (unsigned long long)pg[0], (unsigned long long)pg[1], (long long)sel_target,
                 (long long)(tot_target - sel_target));
    ix->n_sel = (int64_t)pc[0];
    ix->n_next = (int64_t)pc[1];
    if (n_selected) *n_selected = ix->n_sel;
    if (n_next) *n_next = ix->n_next;
    return ANNB_OK;
}

ANNB_API int annb_index_get_thresh(annb_index *ix, double *thresh)
{
    ANNB_REQUIRE(ix && thresh, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(ix->have_thresh, ANNB_ESTATE, "thresholds have not been computed yet");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    std::vector<float> h(ix->n);
    ANNB_CUDA(cudaMemcpyAsync(h.data(), ix->thresh.p, (size_t)ix->n * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t i = 0; i < ix->n; ++i) thresh[i] = h[i];
    return ANNB_OK;
}

// load a caller-chosen look-ahead set (what annb_index_select leaves behind as `next`), so that
// annb_index_update_bounds can be driven and checked on explicit pairs
ANNB_API int annb_index_set_lookahead(annb_index *ix, const int64_t *ij, int64_t m)
{
    TraceScope _ts("annb_index_set_lookahead");
    ANNB_REQUIRE(ix && (m == 0 || ij), ANNB_EINVAL, "NULL argument");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    ix->n_next = 0;
    if (m == 0) return ANNB_OK;
    ANNB_TRY(upload_pairs(ix, ij, nullptr, m));
    ANNB_TRY(ix->nxt_i.ensure((size_t)m * 4));
    ANNB_TRY(ix->nxt_j.ensure((size_t)m * 4));
    ANNB_CUDA(cudaMemcpyAsync(ix->nxt_i.p, ix->t1.p, (size_t)m * 4, cudaMemcpyDeviceToDevice, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(ix->nxt_j.p, ix->t2.p, (size_t)m * 4, cudaMemcpyDeviceToDevice, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    ix->n_next = m;
    return ANNB_OK;
}

ANNB_API int annb_index_get_selected(annb_index *ix, int64_t *ij_sel, int64_t *ij_next)
{
    TraceScope _ts("annb_index_get_selected");
    ANNB_REQUIRE(ix != nullptr, ANNB_EINVAL, "index is NULL");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    auto fetch = [&](const DevBuf &bi, const DevBuf &bj, int64_t m, int64_t *out) -> int {
        if (!out || m == 0) return ANNB_OK;
        std::vector<int32_t> hi(m), hj(m);
        ANNB_CUDA(cudaMemcpyAsync(hi.data(), bi.p, (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaMemcpyAsync(hj.data(), bj.p, (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
        for (int64_t p = 0; p < m; ++p) {
            out[2 * p] = hi[p];
            out[2 * p + 1] = hj[p];
        }
        return ANNB_OK;
    };
    ANNB_TRY(fetch(ix->sel_i, ix->sel_j, ix->n_sel, ij_sel));
    ANNB_TRY(fetch(ix->nxt_i, ix->nxt_j, ix->n_next, ij_next));
    return ANNB_OK;
}

ANNB_API int annb_index_refine_selected(annb_index *ix, int64_t *n_evals)
{
    TraceScope _ts("annb_index_refine_selected");
    ANNB_REQUIRE(ix != nullptr, ANNB_EINVAL, "index is NULL");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    const int64_t m = ix->n_sel;
    if (n_evals) *n_evals = m;
    if (m) {
        ANNB_TRY(ix->t3.ensure((size_t)m * 4));
        ANNB_TRY(pair_dists_f32_perm(c, ix->ds, ix->metric, ix->sel_i.as<int32_t>(),
                                     ix->sel_j.as<int32_t>(), nullptr, m, ix->t3.as<float>()));
        ANNB_TRY(hash_insert(ix, ix->sel_i.as<int32_t>(), ix->sel_j.as<int32_t>(), ix->t3.as<float>(),
                             nullptr, KIND_KNOWN, m));
        ix->n_known += m;
    }
    ix->n_refined = m;  // (sel_i, sel_j, t3) stay valid until the next select for annb_index_export_refined
    if (ix->has_forced) {
        // the -1 marks of guarantee_nmin do not survive the iteration (annchor.py:374-379)
        ANNB_LAUNCH(hash_retire_forced_kernel, grid_for_n(c, (int64_t)ix->hcap), 256, 0, c->stream,
                    ix->htab.as<HashSlot>(), ix->hcap);
        ix->has_forced = false;
        ix->tl_dirty = true;
    }
    ix->n_sel = 0;
    return ANNB_OK;
}

static int build_known_csr(annb_index *ix)
{
    annb_ctx *c = ix->ctx;
    const int64_t n = ix->n;
    ANNB_TRY(ix->kdeg.ensure((size_t)(n + 1) * 4));
    ANNB_TRY(ix->kptr.ensure((size_t)(n + 1) * 8));
    ANNB_CUDA(cudaMemsetAsync(ix->kdeg.p, 0, (size_t)(n + 1) * 4, c->stream));
    ANNB_LAUNCH(known_degree_kernel, grid_for_n(c, (int64_t)ix->hcap), 256, 0, c->stream,
                ix->htab.as<HashSlot>(), ix->hcap, ix->kdeg.as<int32_t>());
    ANNB_TRY(launch_scan_i32_i64(c, ix->kdeg.as<int32_t>(), ix->kptr.as<int64_t>(), n, ix->scan_tmp));
    int64_t total = 0;
    ANNB_CUDA(cudaMemcpyAsync(&total, ix->kptr.as<int64_t>() + n, 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    ix->csr_entries = total;
    ANNB_TRY(ix->kids.ensure((size_t)(total + 1) * 4));
    ANNB_TRY(ix->kds.ensure((size_t)(total + 1) * 4));
    ANNB_CUDA(cudaMemsetAsync(ix->kdeg.p, 0, (size_t)(n + 1) * 4, c->stream));
    ANNB_LAUNCH(known_fill_kernel, grid_for_n(c, (int64_t)ix->hcap), 256, 0, c->stream,
                ix->htab.as<HashSlot>(), ix->hcap, ix->kptr.as<int64_t>(), ix->kdeg.as<int32_t>(),
                ix->kids.as<int32_t>(), ix->kds.as<float>());
    // rows stay unsorted: the tightening kernel hashes / streams them and the top-k is order-free
    return ANNB_OK;
}

ANNB_API int annb_index_update_bounds(annb_index *ix, int64_t *n_updated)
{
    TraceScope _ts("annb_index_update_bounds");
    ANNB_REQUIRE(ix != nullptr, ANNB_EINVAL, "index is NULL");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    const int64_t m = ix->n_next;
    if (n_updated) *n_updated = 0;
    ix->n_tightened = 0;
    if (m == 0) return ANNB_OK;
    ANNB_TRY(build_known_csr(ix));
    ANNB_TRY(ix->t0.ensure((size_t)m * 4));
    ANNB_TRY(ix->t1.ensure((size_t)m * 4));
    ANNB_TRY(ix->t2.ensure((size_t)m));
    ANNB_TRY(ix->t3.ensure((size_t)m * 4));
    ANNB_TRY(ix->t4.ensure((size_t)m * 4));
    ANNB_TRY(ix->t5.ensure((size_t)m * 4));
    ANNB_TRY(ix->t6.ensure((size_t)m * 4));
    const View V = ix->view();
    // group the look-ahead pairs by their lower endpoint
    const int64_t n = ix->n;
    ANNB_TRY(ix->kdeg.ensure((size_t)(n + 1) * 4));
    ANNB_TRY(ix->gptr.ensure((size_t)(n + 1) * 8));
    ANNB_TRY(ix->gJ.ensure((size_t)m * 4));
    ANNB_TRY(ix->gsrc.ensure((size_t)m * 4));
    ANNB_CUDA(cudaMemsetAsync(ix->kdeg.p, 0, (size_t)(n + 1) * 4, c->stream));
    ANNB_TRY(ix->twork.ensure((size_t)n * 8));
    ANNB_TRY(ix->theavy.ensure((size_t)n * 4));
    ANNB_CUDA(cudaMemsetAsync(ix->twork.p, 0, (size_t)n * 8, c->stream));
    ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
    ANNB_LAUNCH(count_by_lo_kernel, grid_for_n(c, m), 256, 0, c->stream, ix->nxt_i.as<int32_t>(),
                ix->nxt_j.as<int32_t>(), m, ix->kptr.as<int64_t>(), ix->kdeg.as<int32_t>(),
                ix->twork.as<unsigned long long>(), ix->counters.as<unsigned long long>() + 7);
    unsigned long long wsum[3] = {0, 0, 0};
    ANNB_CUDA(cudaMemcpyAsync(wsum, ix->counters.as<unsigned long long>() + 7, 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_TRY(launch_scan_i32_i64(c, ix->kdeg.as<int32_t>(), ix->gptr.as<int64_t>(), n, ix->scan_tmp));
    ANNB_CUDA(cudaMemsetAsync(ix->kdeg.p, 0, (size_t)(n + 1) * 4, c->stream));
    ANNB_LAUNCH(scatter_by_lo_kernel, grid_for_n(c, m), 256, 0, c->stream, ix->nxt_i.as<int32_t>(),
                ix->nxt_j.as<int32_t>(), m, ix->kptr.as<int64_t>(), ix->gptr.as<int64_t>(),
                ix->kdeg.as<int32_t>(), ix->gJ.as<int32_t>(), ix->gsrc.as<int32_t>());
    if (g_debug_sync) {
        unsigned long long w[3] = {0, 0, 0};
        ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
        ANNB_TRY(ix->t5.ensure((size_t)n * 8));
        ANNB_LAUNCH(tighten_work_kernel, grid_for_n(c, n), 256, 0, c->stream, ix->kptr.as<int64_t>(),
                    ix->gptr.as<int64_t>(), ix->gJ.as<int32_t>(), n, ix->counters.as<unsigned long long>(),
                    ix->t5.as<unsigned long long>());
        std::vector<unsigned long long> rw(n);
        std::vector<int64_t> kp(n + 1), gp(n + 1);
        ANNB_CUDA(cudaMemcpyAsync(w, ix->counters.p, 24, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaMemcpyAsync(rw.data(), ix->t5.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaMemcpyAsync(kp.data(), ix->kptr.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaMemcpyAsync(gp.data(), ix->gptr.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
        fprintf(stderr, "[annb] tighten: %lld pairs, sum deg_j = %llu (%.1f/pair), sum min(deg) = %llu, max deg = %llu, csr = %lld\n",
                (long long)m, w[0], (double)w[0] / (double)m, w[1], w[2], (long long)ix->csr_entries);
        std::vector<int64_t> deg(n), gs(n);
        for (int64_t i = 0; i < n; ++i) {
            deg[i] = kp[i + 1] - kp[i];
            gs[i] = gp[i + 1] - gp[i];
        }
        auto q = [&](std::vector<int64_t> v, const char *nm) {
            std::sort(v.begin(), v.end());
            fprintf(stderr, "[annb]   %s: p50 %lld p90 %lld p99 %lld p99.9 %lld max %lld\n", nm, (long long)v[n / 2],
                    (long long)v[n * 9 / 10], (long long)v[n * 99 / 100], (long long)v[n * 999 / 1000], (long long)v[n - 1]);
        };
        q(deg, "known degree");
        q(gs, "group size  ");
        std::vector<int64_t> rwi(rw.begin(), rw.end());
        q(rwi, "row work    ");
        // imbalance of the static round-robin row -> CTA assignment
        const int G = c->num_sms * 3;
        std::vector<unsigned long long> cta(G, 0);
        for (int64_t r = 0; r < n; ++r) cta[r % G] += rw[r];  // (ignores row_order; same statistics)
        unsigned long long mx = 0, sm = 0;
        for (int g = 0; g < G; ++g) {
            mx = std::max(mx, cta[g]);
            sm += cta[g];
        }
        fprintf(stderr, "[annb]   CTA work: max/mean = %.2f\n", (double)mx / ((double)sm / G));
    }
    if (ix->row_order.p == nullptr) {
        ANNB_TRY(ix->row_order.ensure((size_t)n * 4));
        ANNB_TRY(ix->t5.ensure((size_t)std::max<int64_t>(m * 4, (ROW_BUCKETS + 1) * 16)));  // hist | ptr, reused below
        int32_t *rb_hist = ix->t5.as<int32_t>();
        int64_t *rb_ptr = reinterpret_cast<int64_t *>(rb_hist + ROW_BUCKETS + 2);
        // scale of the distance buckets: the largest anchor distance
        ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
        ANNB_LAUNCH(max_f32_kernel, grid_for_n(c, (int64_t)ix->na * ix->npad), 256, 0, c->stream,
                    ix->D32.as<float>(), (int64_t)ix->na * ix->npad, ix->counters.as<unsigned int>());
        float dmax = 0.0f;
        ANNB_CUDA(cudaMemcpyAsync(&dmax, ix->counters.p, 4, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
        const float inv_scale = dmax > 0.0f && std::isfinite(dmax) ? 1023.0f / dmax : 0.0f;
        ANNB_CUDA(cudaMemsetAsync(rb_hist, 0, (size_t)(ROW_BUCKETS + 2) * 4, c->stream));
        ANNB_LAUNCH(row_bucket_hist_kernel, grid_for_n(c, n), 256, 0, c->stream, ix->meta.as<PointMeta>(),
                    ix->D32.as<float>(), ix->npad, n, inv_scale, rb_hist);
        ANNB_TRY(launch_scan_i32_i64(c, rb_hist, rb_ptr, ROW_BUCKETS, ix->scan_tmp));
        ANNB_CUDA(cudaMemsetAsync(rb_hist, 0, (size_t)(ROW_BUCKETS + 2) * 4, c->stream));
        ANNB_LAUNCH(row_bucket_scatter_kernel, grid_for_n(c, n), 256, 0, c->stream, ix->meta.as<PointMeta>(),
                    ix->D32.as<float>(), ix->npad, n, inv_scale, rb_ptr, rb_hist, ix->row_order.as<int32_t>());
    }
    // per-row streaming work (from the grouping pass) -> heavy rows first, then closest-anchor order
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    const int tg_grid = c->num_sms * 3;
    // a row is "heavy" when it is more than 1/16 of a CTA's fair share of the streamed entries
    const unsigned long long heavy_thr = std::max<unsigned long long>(wsum[0] / ((unsigned long long)tg_grid * 16), 1);
    ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
    ANNB_LAUNCH(tighten_heavy_kernel, grid_for_n(c, n), 256, 0, c->stream, ix->twork.as<unsigned long long>(), n,
                heavy_thr, ix->theavy.as<int32_t>(), ix->counters.as<unsigned long long>() + 4);
    // membership-bitmap kernel while a bit per point fits shared memory (N <~ 1.2 M), else the bucket hash
    const int W = (int)((n + 1023) / 1024 * 32);
    const size_t tb_smem = (size_t)W * 4 + TB_CHUNK * 4 + (size_t)W * 2 + 64;
    const bool force_hash = getenv("ANNB_TIGHTEN_HASH") != nullptr;  // test knob: bucket-hash kernel
    int tb_chunk = getenv("ANNB_TB_CHUNK") ? atoi(getenv("ANNB_TB_CHUNK")) : TB_CHUNK;  // test knob: small chunks
    tb_chunk = std::max(32, std::min(tb_chunk, TB_CHUNK));
    if (tb_smem <= 220 * 1024 && !force_hash) {
        // as many 512-thread CTAs per SM as the table allows (<= 4); one 1024-thread CTA for large N
        const int per_sm = (int)std::min<size_t>(4, (227 * 1024) / (tb_smem + 1024));
        if (per_sm >= 2) {
            ANNB_CUDA(cudaFuncSetAttribute(tighten_bitmap_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)tb_smem));
            ANNB_LAUNCH(tighten_bitmap_kernel<512>, c->num_sms * per_sm, 512, tb_smem, c->stream, V,
                        ix->kptr.as<int64_t>(), ix->kids.as<int32_t>(), ix->kds.as<float>(), ix->gptr.as<int64_t>(),
                        ix->gJ.as<int32_t>(), ix->gsrc.as<int32_t>(), ix->row_order.as<int32_t>(),
                        ix->theavy.as<int32_t>(), ix->counters.as<unsigned long long>() + 4,
                        ix->twork.as<unsigned long long>(), heavy_thr, ix->n_tight > 0 ? 1 : 0, ix->t0.as<float>(),
                        ix->t1.as<float>(), ix->t2.as<uint8_t>(), W, tb_chunk);
        } else {
            ANNB_CUDA(cudaFuncSetAttribute(tighten_bitmap_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)tb_smem));
            ANNB_LAUNCH(tighten_bitmap_kernel<1024>, c->num_sms, 1024, tb_smem, c->stream, V,
                        ix->kptr.as<int64_t>(), ix->kids.as<int32_t>(), ix->kds.as<float>(), ix->gptr.as<int64_t>(),
                        ix->gJ.as<int32_t>(), ix->gsrc.as<int32_t>(), ix->row_order.as<int32_t>(),
                        ix->theavy.as<int32_t>(), ix->counters.as<unsigned long long>() + 4,
                        ix->twork.as<unsigned long long>(), heavy_thr, ix->n_tight > 0 ? 1 : 0, ix->t0.as<float>(),
                        ix->t1.as<float>(), ix->t2.as<uint8_t>(), W, tb_chunk);
        }
    } else {
        ANNB_CUDA(cudaFuncSetAttribute(tighten_grouped_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       TG_SLOTS * 8));
        ANNB_LAUNCH(tighten_grouped_kernel, tg_grid, 512, TG_SLOTS * 8, c->stream, V, ix->kptr.as<int64_t>(),
                    ix->kids.as<int32_t>(), ix->kds.as<float>(), ix->gptr.as<int64_t>(), ix->gJ.as<int32_t>(),
                    ix->gsrc.as<int32_t>(), ix->row_order.as<int32_t>(), ix->theavy.as<int32_t>(),
                    ix->counters.as<unsigned long long>() + 4, ix->twork.as<unsigned long long>(), heavy_thr,
                    ix->n_tight > 0 ? 1 : 0, ix->t0.as<float>(), ix->t1.as<float>(), ix->t2.as<uint8_t>());
    }
    ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
    ANNB_LAUNCH(compact_improved_kernel, grid_for_n(c, m), 256, 0, c->stream, ix->nxt_i.as<int32_t>(),
                ix->nxt_j.as<int32_t>(), ix->t0.as<float>(), ix->t1.as<float>(), ix->t2.as<uint8_t>(), m,
                ix->t3.as<int32_t>(), ix->t4.as<int32_t>(), ix->t5.as<float>(), ix->t6.as<float>(),
                ix->counters.as<unsigned long long>());
    unsigned long long k = 0;
    ANNB_CUDA(cudaMemcpyAsync(&k, ix->counters.p, 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    ANNB_TRY(hash_insert(ix, ix->t3.as<int32_t>(), ix->t4.as<int32_t>(), ix->t5.as<float>(),
                         ix->t6.as<float>(), KIND_TIGHT, (int64_t)k));
    ix->n_tight += (int64_t)k;
    ix->n_tightened = (int64_t)k;  // (t3, t4, t5, t6) stay valid for annb_index_export_tightened
    if (n_updated) *n_updated = (int64_t)k;
    ix->n_next = 0;
    return ANNB_OK;
}

ANNB_API int annb_index_neighbor_graph(annb_index *ix, int64_t *idx, double *dist)
{
    TraceScope _ts("annb_index_neighbor_graph");
    ANNB_REQUIRE(ix && idx && dist, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(ix->have_anchors, ANNB_ESTATE, "anchors not computed yet");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    const int64_t n = ix->n;
    const int nn = ix->P.n_neighbors;
    ANNB_TRY(build_known_csr(ix));
    ANNB_TRY(ix->t0.ensure((size_t)n * nn * 8));
    ANNB_TRY(ix->t1.ensure((size_t)n * nn * 8));
    ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
    const int grid = (int)std::min<int64_t>(n, (int64_t)c->num_sms * 8);
    ANNB_LAUNCH(neighbor_graph_kernel, grid, 256, 0, c->stream, ix->view(), ix->D64.as<double>(),
                ix->A_dev.as<int32_t>(), (int)ix->A_host.size(), ix->kptr.as<int64_t>(),
                ix->kids.as<int32_t>(), ix->kds.as<float>(), ix->t0.as<int64_t>(), ix->t1.as<double>(),
                ix->counters.as<int32_t>());
    int32_t n_deficient = 0;
    ANNB_CUDA(cudaMemcpyAsync(&n_deficient, ix->counters.p, 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    // a renumbered index (annb_index_adopt_anchors) reports the graph in the caller's numbering
    const int32_t *order = ix->ordered ? ix->order_dev.as<int32_t>() : nullptr;
    if (n_deficient == 0 && order) {
        ANNB_TRY(ix->t2.ensure((size_t)n * nn * 8));
        ANNB_TRY(ix->t3.ensure((size_t)n * nn * 8));
        ANNB_LAUNCH(unpermute_graph_kernel, grid_for_n(c, n * nn), 256, 0, c->stream, ix->t0.as<int64_t>(),
                    ix->t1.as<double>(), n, nn, order, ix->t2.as<int64_t>(), ix->t3.as<double>());
        ANNB_CUDA(cudaMemcpyAsync(idx, ix->t2.p, (size_t)n * nn * 8, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaMemcpyAsync(dist, ix->t3.p, (size_t)n * nn * 8, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
        return ANNB_OK;
    }
    ANNB_CUDA(cudaMemcpyAsync(idx, ix->t0.p, (size_t)n * nn * 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(dist, ix->t1.p, (size_t)n * nn * 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    // (rows that need the prediction fall-back are patched on the host in the index's numbering first)
    auto to_caller = [&]() -> int {
        if (!order) return ANNB_OK;
        std::vector<int32_t> o(n);
        ANNB_CUDA(cudaMemcpy(o.data(), order, (size_t)n * 4, cudaMemcpyDeviceToHost));
        std::vector<int64_t> ti((size_t)n * nn);
        std::vector<double> td((size_t)n * nn);
        for (int64_t r = 0; r < n; ++r)
            for (int q = 0; q < nn; ++q) {
                const int64_t id = idx[r * nn + q];
                ti[(size_t)o[r] * nn + q] = id >= 0 ? o[id] : -1;
                td[(size_t)o[r] * nn + q] = dist[r * nn + q];
            }
        memcpy(idx, ti.data(), ti.size() * 8);
        memcpy(dist, td.data(), td.size() * 8);
        return ANNB_OK;
    };
    if (n_deficient == 0) return to_caller();
    // Rows with fewer than nn-1 computed pairs: the reference's get_nn ranks the not-computed
    // candidates behind the computed ones by their RefineApprox (d[ncm] += max(d), utils.py:415-428)
    // and emits the prediction.  Same here: the row sweep's second list gives, for the affected row
    // blocks, the nn-1 smallest not-computed predictions with their ids.
    ANNB_REQUIRE(ix->have_model, ANNB_ESTATE,
                 "%d neighbour slots have no computed pair and no regression model is set to rank predictions",
                 (int)n_deficient);
    const int k2 = nn - 1;
    std::vector<int32_t> rbs;
    {
        std::vector<char> bad((size_t)ix->T, 0);
        for (int64_t r = 0; r < n; ++r)
            if (idx[r * nn + nn - 1] < 0) bad[r / TILE] = 1;
        for (int t = 0; t < ix->T; ++t)
            if (bad[t]) rbs.push_back(t);
    }
    ANNB_TRY(ix->l2val.ensure((size_t)n * k2 * 4));
    ANNB_TRY(ix->l2id.ensure((size_t)n * k2 * 4));
    ANNB_TRY(ix->t3.ensure(rbs.size() * 4));
    ANNB_CUDA(cudaMemcpyAsync(ix->t3.p, rbs.data(), rbs.size() * 4, cudaMemcpyHostToDevice, c->stream));
    ANNB_TRY(run_thresh_rows(ix, k2, 1, nullptr, ix->t3.as<int32_t>(), (int)rbs.size()));
    ix->have_thresh = false;  // the sweep overwrote the thresholds of those row blocks
    std::vector<float> lv((size_t)n * k2);
    std::vector<int32_t> li((size_t)n * k2);
    ANNB_CUDA(cudaMemcpyAsync(lv.data(), ix->l2val.p, lv.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(li.data(), ix->l2id.p, li.size() * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t r = 0; r < n; ++r) {
        if (idx[r * nn + nn - 1] >= 0) continue;
        int col = 1;
        while (col < nn && idx[r * nn + col] >= 0) ++col;
        for (int q = 0; col < nn && q < k2; ++q) {
            const int32_t id = li[(size_t)r * k2 + q];
            if (id < 0) break;  // fewer candidates than neighbours asked for
            idx[r * nn + col] = id;
            dist[r * nn + col] = (double)lv[(size_t)r * k2 + q];
            ++col;
        }
    }
    return to_caller();
}

ANNB_API int annb_index_stats(annb_index *ix, int64_t *out, int64_t m)
{
    TraceScope _ts("annb_index_stats");
    ANNB_REQUIRE(ix && out, ANNB_EINVAL, "NULL argument");
    const int64_t v[8] = {ix->pairs_swept, ix->n_known,        ix->n_tight,        ix->sweeps,
                          ix->n_candidates, (int64_t)ix->hcap, ix->n_anchor_pairs, ix->n_not_computed()};
    for (int64_t k = 0; k < m && k < 8; ++k) out[k] = v[k];
    return ANNB_OK;
}

ANNB_API int annb_index_last_sweep(annb_index *ix, float *ms, int64_t *pairs)
{
    TraceScope _ts("annb_index_last_sweep");
    ANNB_REQUIRE(ix != nullptr, ANNB_EINVAL, "index is NULL");
    if (ms) *ms = ix->last_sweep_ms;
    if (pairs) *pairs = ix->last_sweep_pairs;
    return ANNB_OK;
}

// ---------------------------------------------------------------------------------------------
// sampler support (annchor/samplers.py:75-140, annchor/utils.py:543-578): a uniform pool of the
// not-computed candidate pairs with their dad.  When there are at most max_pool such pairs the
// pool is ALL of them, in any order (the host sorts by (i, j) = the reference's IJs order and can
// then reproduce the reference's sampler exactly); otherwise a hash-selected uniform sub-sample.
// ---------------------------------------------------------------------------------------------
ANNB_API int annb_index_sample_pool(annb_index *ix, uint64_t seed, int64_t max_pool, int64_t *n_pool,
                                    int64_t *n_not_computed, int *exact)
{
    TraceScope _ts("annb_index_sample_pool");
    ANNB_REQUIRE(ix && n_pool, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(ix->have_locality, ANNB_ESTATE, "locality must run before sampling");
    ANNB_REQUIRE(max_pool >= 1024, ANNB_EINVAL, "max_pool too small");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    const int64_t n_nc = ix->n_not_computed();
    if (n_not_computed) *n_not_computed = n_nc;
    const bool all = n_nc <= max_pool;
    if (exact) *exact = all ? 1 : 0;
    ix->n_pool = 0;
    *n_pool = 0;
    if (n_nc <= 0) return ANNB_OK;
    SampleArgs A;
    A.V = ix->view();
    A.nb = 0;
    A.seed = (uint32_t)(mix64(seed) >> 32);
    double frac = all ? 1.0 : 0.8 * (double)max_pool / (double)n_nc;
    for (int attempt = 0;; ++attempt) {
        ANNB_REQUIRE(attempt < 4, ANNB_ESTATE, "sampler pool did not fit");
        // sparse samples visit a fraction of the tiles and take pairs there at a correspondingly higher rate: about 16
        // pairs per visited tile, between 1 tile in 64 and 1 in 4096 (N = 1M: 2*10^4 of 3*10^7 tiles)
        const double tile_frac = frac < 1.0 / 8192.0 ? std::min(1.0 / 64.0, std::max(1.0 / 4096.0, 1024.0 * frac)) : 1.0;
        A.tile_thr = tile_frac >= 1.0 ? 0xffffffffu : (uint32_t)(tile_frac * 4294967295.0);
        const double pfrac = frac / tile_frac;
        A.thr = pfrac >= 1.0 ? 0xffffffffu : (uint32_t)(pfrac * 4294967295.0);
        const int64_t cap = all ? n_nc + 1024 : max_pool + max_pool / 4;
        ANNB_TRY(ix->pool_key.ensure((size_t)cap * 8));
        ANNB_TRY(ix->pool_dad.ensure((size_t)cap * 4));
        A.out_key = ix->pool_key.as<uint64_t>();
        A.out_dad = ix->pool_dad.as<float>();
        A.out_cap = (unsigned long long)cap;
        ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
        A.counter = ix->counters.as<unsigned long long>();
        A.rank = 0;  // the pool is cheap and replicated: every rank sweeps every tile
        A.world = 1;
        ANNB_TRY(launch_sample_sweep(c, A));
        ix->sweeps += 1;
        unsigned long long got = 0;
        ANNB_CUDA(cudaMemcpyAsync(&got, ix->counters.p, 8, cudaMemcpyDeviceToHost, c->stream));
        ANNB_CUDA(cudaStreamSynchronize(c->stream));
        if ((int64_t)got <= cap) {
            ix->n_pool = (int64_t)got;
            break;
        }
        ANNB_REQUIRE(!all, ANNB_ESTATE, "pool sweep found %llu pairs but %lld were expected", got,
                     (long long)n_nc);
        frac *= 0.7 * (double)cap / (double)got;
    }
    if (all)
        ANNB_REQUIRE(ix->n_pool == n_nc, ANNB_ESTATE,
                     "pool sweep found %lld not-computed candidates, bookkeeping says %lld",
                     (long long)ix->n_pool, (long long)n_nc);
    *n_pool = ix->n_pool;
    return ANNB_OK;
}

// Stratified refill of the pool: a bin of the sampler's partition that the uniform pool left (nearly)
// empty is re-sampled at its own rate over ALL tiles, so that rare dad ranges are represented the way
// the reference's per-bin draw over the materialised pair list represents them.
ANNB_API int annb_index_sample_pool_bins(annb_index *ix, uint64_t seed, const double *bins, const double *rate,
                                         int64_t nb, int64_t max_pool, int64_t *n_pool)
{
    TraceScope _ts("annb_index_sample_pool_bins");
    ANNB_REQUIRE(ix && bins && rate && n_pool, ANNB_EINVAL, "NULL argument");
    ANNB_REQUIRE(ix->have_locality, ANNB_ESTATE, "locality must run before sampling");
    ANNB_REQUIRE(nb >= 1 && nb <= MAX_BINS && max_pool >= 1024, ANNB_ERANGE, "bad bin count / pool size");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    SampleArgs A;
    A.V = ix->view();
    A.seed = (uint32_t)(mix64(seed ^ 0x5851f42d4c957f2dull) >> 32);
    A.nb = (int)nb;
    for (int b = 0; b <= nb; ++b) A.edge[b] = (float)bins[b];
    // sparse rates: visit a fraction of the tiles and take pairs there at a higher rate (like the uniform pool), so
    // that the sweep's cost follows the largest rate instead of being a full pass over all pairs
    double rmax = 0.0;
    for (int b = 0; b < nb; ++b) rmax = std::max(rmax, rate[b]);
    // (in-tile rate <= 1/2) ... and at most ~2.5*10^5 tiles are visited (bounded cost at N = 1M, where a full pass is
    // 3*10^7 tiles): a very rare bin then gets fewer pairs than asked for, still far more than the uniform pool held
    double tile_frac = std::min(1.0, std::max(1.0 / 4096.0, 2.0 * rmax));
    tile_frac = std::min(tile_frac, std::max(1.0 / 4096.0, 2.5e5 / (double)ix->NT));
    A.tile_thr = tile_frac >= 1.0 ? 0xffffffffu : (uint32_t)(tile_frac * 4294967295.0);
    A.thr = 0;
    for (int b = 0; b < MAX_BINS; ++b) {
        const double r = (b < nb ? rate[b] : 0.0) / tile_frac;
        A.bthr[b] = r <= 0.0 ? 0u : (r >= 1.0 ? 0xffffffffu : (uint32_t)std::max(1.0, r * 4294967295.0));
        A.thr = std::max(A.thr, A.bthr[b]);
    }
    ANNB_TRY(ix->pool_key.ensure((size_t)max_pool * 8));
    ANNB_TRY(ix->pool_dad.ensure((size_t)max_pool * 4));
    A.out_key = ix->pool_key.as<uint64_t>();
    A.out_dad = ix->pool_dad.as<float>();
    A.out_cap = (unsigned long long)max_pool;
    ANNB_CUDA(cudaMemsetAsync(ix->counters.p, 0, 64, c->stream));
    A.counter = ix->counters.as<unsigned long long>();
    A.rank = 0;
    A.world = 1;
    ANNB_TRY(launch_sample_sweep(c, A));
    ix->sweeps += 1;
    unsigned long long got = 0;
    ANNB_CUDA(cudaMemcpyAsync(&got, ix->counters.p, 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    ix->n_pool = (int64_t)std::min<unsigned long long>(got, (unsigned long long)max_pool);
    *n_pool = ix->n_pool;
    return ANNB_OK;
}

ANNB_API int annb_index_get_pool(annb_index *ix, int64_t *ij, double *dad)
{
    TraceScope _ts("annb_index_get_pool");
    ANNB_REQUIRE(ix && ij && dad, ANNB_EINVAL, "NULL argument");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    const int64_t m = ix->n_pool;
    if (m == 0) return ANNB_OK;
    std::vector<uint64_t> k(m);
    std::vector<float> d(m);
    ANNB_CUDA(cudaMemcpyAsync(k.data(), ix->pool_key.p, (size_t)m * 8, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(d.data(), ix->pool_dad.p, (size_t)m * 4, cudaMemcpyDeviceToHost, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    for (int64_t p = 0; p < m; ++p) {
        ij[2 * p] = (int64_t)(k[p] >> 32);
        ij[2 * p + 1] = (int64_t)(k[p] & 0xffffffffu);
        dad[p] = d[p];
    }
    return ANNB_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU plumbing: reducer hook + device-to-device export / import of per-rank results
// ---------------------------------------------------------------------------------------------
ANNB_API int annb_index_set_reducer(annb_index *ix, annb_reduce_fn fn, void *user)
{
    TraceScope _ts("annb_index_set_reducer");
    ANNB_REQUIRE(ix != nullptr, ANNB_EINVAL, "index is NULL");
    ix->reducer = fn;
    ix->reducer_user = user;
    return ANNB_OK;
}

ANNB_API int annb_index_export_refined(annb_index *ix, int32_t *i_dev, int32_t *j_dev, float *d_dev,
                                       int64_t cap, int64_t *n)
{
    TraceScope _ts("annb_index_export_refined");
    ANNB_REQUIRE(ix && n, ANNB_EINVAL, "NULL argument");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    *n = ix->n_refined;
    if (ix->n_refined == 0 || !i_dev) return ANNB_OK;
    ANNB_REQUIRE(cap >= ix->n_refined, ANNB_ERANGE, "export buffer too small");
    const size_t b = (size_t)ix->n_refined * 4;
    ANNB_CUDA(cudaMemcpyAsync(i_dev, ix->sel_i.p, b, cudaMemcpyDeviceToDevice, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(j_dev, ix->sel_j.p, b, cudaMemcpyDeviceToDevice, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(d_dev, ix->t3.p, b, cudaMemcpyDeviceToDevice, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    return ANNB_OK;
}

ANNB_API int annb_index_export_tightened(annb_index *ix, int32_t *i_dev, int32_t *j_dev, float *lb_dev,
                                         float *ub_dev, int64_t cap, int64_t *n)
{
    TraceScope _ts("annb_index_export_tightened");
    ANNB_REQUIRE(ix && n, ANNB_EINVAL, "NULL argument");
    annb_ctx *c = ix->ctx;
    ANNB_CUDA(cudaSetDevice(c->device));
    *n = ix->n_tightened;
    if (ix->n_tightened == 0 || !i_dev) return ANNB_OK;
    ANNB_REQUIRE(cap >= ix->n_tightened, ANNB_ERANGE, "export buffer too small");
    const size_t b = (size_t)ix->n_tightened * 4;
    ANNB_CUDA(cudaMemcpyAsync(i_dev, ix->t3.p, b, cudaMemcpyDeviceToDevice, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(j_dev, ix->t4.p, b, cudaMemcpyDeviceToDevice, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(lb_dev, ix->t5.p, b, cudaMemcpyDeviceToDevice, c->stream));
    ANNB_CUDA(cudaMemcpyAsync(ub_dev, ix->t6.p, b, cudaMemcpyDeviceToDevice, c->stream));
    ANNB_CUDA(cudaStreamSynchronize(c->stream));
    return ANNB_OK;
}

// kind: 1 = exactly known (a = distance), 2 = tightened (a, b) = (lb, ub); all pointers device
ANNB_API int annb_index_import_dev(annb_index *ix, int kind, const int32_t *i_dev, const int32_t *j_dev,
                                   const float *a_dev, const float *b_dev, int64_t n)
{
    TraceScope _ts("annb_index_import_dev");
    ANNB_REQUIRE(ix != nullptr, ANNB_EINVAL, "index is NULL");
    ANNB_REQUIRE(kind == (int)KIND_KNOWN || kind == (int)KIND_TIGHT, ANNB_EINVAL, "kind must be 1 or 2");
    if (n == 0) return ANNB_OK;
    ANNB_REQUIRE(i_dev && j_dev && a_dev, ANNB_EINVAL, "NULL buffer");
    ANNB_CUDA(cudaSetDevice(ix->ctx->device));
    ANNB_TRY(hash_insert(ix, i_dev, j_dev, a_dev, kind == (int)KIND_TIGHT ? b_dev : nullptr, (uint32_t)kind, n));
    if (kind == (int)KIND_KNOWN) ix->n_known += n;
    else ix->n_tight += n;
    ANNB_CUDA(cudaStreamSynchronize(ix->ctx->stream));
    return ANNB_OK;
}
