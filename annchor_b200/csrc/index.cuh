// index.cuh -- device-side data model of the streaming index (annb_index).
//
// The reference materialises, per candidate pair: IJs (16 B), features (32 B), pred, RefineApprox,
// errors, masks (annchor/annchor.py:258-303,345-393) -- Theta(N^2) host memory.  Here the only
// per-pair state is what was actually measured or tightened:
//   * a hash map  pair -> {exact distance | tightened (lb, ub) | forced}, 16 B per slot   (HBM)
//   * the same entries listed per 128x128 tile (12 B each), which is what the sweeps read  (HBM)
// Everything else (bounds, dad, prediction, label, probability) is re-derived inside the tile
// sweeps from the anchor-distance matrix D (n_anchors x N float32, anchor-major) staged in
// shared memory.
#pragma once
#include "common.cuh"

namespace annb {

constexpr int TILE = 128;        // points per tile side
constexpr int SROW = 132;        // shared-memory row stride in floats (128 + 4: keeps 16 B alignment,
                                 // spreads the dad look-ups D[i][cA[j]] over banks)
constexpr int BITMAP_WORDS = TILE * TILE / 32;  // 512 words = 2 KB: a tile's flag bitmap (shared memory only)
constexpr int MAX_BINS = 8;
constexpr int MAX_LIST = 64;     // longest per-row sorted list the thresh sweep keeps

// kind of a hash-map entry, stored in the two spare bits (63, 31) of the 64-bit pair key
enum : uint32_t { KIND_NONE = 0, KIND_KNOWN = 1, KIND_TIGHT = 2, KIND_FORCED = 3 };
constexpr uint64_t HKEY_EMPTY = ~0ull;
constexpr uint64_t HKEY_MASK = 0x7fffffff7fffffffull;

struct __align__(16) HashSlot {
    uint64_t key;  // pair key | kind bits; HKEY_EMPTY when free
    float a, b;    // KNOWN: a = distance; TIGHT: (a, b) = (lb, ub)
};

__host__ __device__ __forceinline__ uint64_t kind_bits(uint32_t kind)
{
    return ((uint64_t)(kind >> 1) << 63) | ((uint64_t)(kind & 1u) << 31);
}
__host__ __device__ __forceinline__ uint32_t kind_of(uint64_t stored)
{
    return (uint32_t)(((stored >> 63) << 1) | ((stored >> 31) & 1ull));
}

struct __align__(16) PointMeta {
    uint64_t amask;   // bit a set <=> anchor a is among the `locality` nearest anchors (annchor.py:235-241)
    int32_t cA;       // closest anchor (np.argmin, utils.py:375)
    int8_t loc_t;     // per-row locality threshold min(loc_thresh, kth) (utils.py:472-480)
    int8_t slot;      // position in A if the point is an anchor, else -1
    int16_t pad;
};

struct Model {
    int nb;
    float edge[MAX_BINS + 1];  // edge[0] = -inf, edge[nb] = +inf (annchor/samplers.py:138-139)
    float e2[MAX_BINS];        // 2 * edge[k] for the interior edges k = 1..nb-1, +inf otherwise (sweeps
                               // compare against 2 * dad)
    float c0[MAX_BINS], c1[MAX_BINS], c2[MAX_BINS], ic[MAX_BINS];  // regressors.py:39-67
    int eoff[MAX_BINS + 1];    // offsets of the per-label sorted error tables (error_predictors.py:47-54)
};

struct View {
    int64_t n, npad;
    int na, T, nn, is_metric;
    const float *D32;          // (na, npad) anchor-major
    const float *Dpm;          // (npad, dpitch) point-major copy (one coalesced row per point)
    int dpitch;
    const PointMeta *meta;     // npad
    const HashSlot *htab;
    uint64_t hmask;            // capacity - 1
    // per-tile lists of the store's entries (rebuilt from the hash map after every insert batch)
    const long long *tl_ptr;   // [NT + 1] upper-triangular tiles, row-major (tile_index)
    const uint32_t *tl_code;   // r << 9 | c << 2 | kind
    const float *tl_a, *tl_b;  // KNOWN: a = distance; TIGHT: (a, b) = (lb, ub)
    // per-tile bounds: [T][kMaxAnchors] min / max anchor distance over the tile's points, and the set of
    // closest anchors that occur in the tile -- a sweep skips a tile pair when they prove that no pair of
    // it can pass the phase-1 test (sweep.cuh: tile_can_pass)
    const float *tb_lo, *tb_hi;
    const uint64_t *tb_cm;
    int cull;
};

__host__ __device__ __forceinline__ int64_t tile_index(int ti, int tj, int T)
{
    // ti <= tj; row-major over the upper triangle
    return (int64_t)ti * T - (int64_t)ti * (ti - 1) / 2 + (tj - ti);
}

__host__ __device__ __forceinline__ uint64_t pair_key(uint32_t lo, uint32_t hi)
{
    return ((uint64_t)lo << 32) | hi;
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x)
{
    // splitmix64 finaliser: a bijection on 64-bit keys
    x ^= x >> 30;
    x *= 0xbf58476d1ce4e5b9ull;
    x ^= x >> 27;
    x *= 0x94d049bb133111ebull;
    x ^= x >> 31;
    return x;
}

// 32-bit pair hash (cheap enough for phase 1 of the sweeps)
__host__ __device__ __forceinline__ uint32_t hash_pair32(uint32_t i, uint32_t j, uint32_t seed)
{
    uint32_t h = (i * 0x9E3779B1u) ^ (j * 0x85EBCA77u) ^ seed;
    h ^= h >> 15;
    h *= 0x2C1B3C6Du;
    h ^= h >> 12;
    h *= 0x297A2D39u;
    h ^= h >> 15;
    return h;
}

// tie-break key of a pair at a selection cut (smaller wins): the top 32 bits are the cheap pair hash,
// so phase 1 can pre-filter on them; the low 32 bits make the order total
__host__ __device__ __forceinline__ uint64_t tie_key(uint32_t lo, uint32_t hi, uint64_t salt)
{
    return ((uint64_t)hash_pair32(lo, hi, (uint32_t)salt) << 32) | (uint32_t)mix64(pair_key(lo, hi) ^ salt);
}

// one 16-byte load per probe; returns the kind (KIND_NONE if absent) and fills (a, b)
__device__ __forceinline__ uint32_t hash_lookup(const View &V, uint64_t key, float &a, float &b)
{
    uint64_t h = mix64(key) & V.hmask;
    for (;;) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(V.htab + h));
        const uint64_t k = ((uint64_t)raw.y << 32) | raw.x;
        if (k == HKEY_EMPTY) return KIND_NONE;
        if ((k & HKEY_MASK) == key) {
            a = __uint_as_float(raw.z);
            b = __uint_as_float(raw.w);
            return kind_of(k);
        }
        h = (h + 1) & V.hmask;
    }
}

// regression bin: (lo, hi]  (regressors.py:85-87) -> number of interior edges strictly below dad
__device__ __forceinline__ int reg_bin(const Model &M, float dad)
{
    int b = 0;
#pragma unroll
    for (int k = 1; k < MAX_BINS; ++k) b += (k < M.nb && dad > M.edge[k]);
    return b;
}
// error label: closed [lo, hi], later bins win (error_predictors.py:63-66); the sampler's
// [lo, hi) bins (utils.py:547-549) give the same index
__device__ __forceinline__ int err_label(const Model &M, float dad)
{
    int b = 0;
#pragma unroll
    for (int k = 1; k < MAX_BINS; ++k) b += (k < M.nb && dad >= M.edge[k]);
    return b;
}

__device__ __forceinline__ float predict_clip(const Model &M, float lb, float ub, float dad)
{
    const int b = reg_bin(M, dad);
    const float y = lb * M.c0[b] + ub * M.c1[b] + dad * M.c2[b] + M.ic[b];
    return fminf(fmaxf(y, lb), ub);  // np.clip(pred, lb, ub), annchor.py:361-363
}

}  // namespace annb
