// sweep.cuh -- building blocks shared by the tile-sweep kernels (thresh / score / sample).
//
// A tile is 128 x 128 pairs handled by one 256-thread CTA; each thread owns an 8 x 8 micro-tile:
// rows {4*ty..4*ty+3} U {64+4*ty..}, columns {4*tx..4*tx+3} U {64+4*tx..}, (ty, tx) = (tid/16, tid%16),
// so every shared-memory read of anchor-distance rows is a conflict-free 128-bit access
// (8 consecutive threads cover 128 contiguous bytes; the two ty values in a warp broadcast).
// A warp therefore owns 16 whole rows of the tile ({8w..8w+7} U {64+8w..64+8w+7}) x all 128 columns.
//
// Two-phase tile processing keeps the hot loop tiny:
//   phase 1 (every pair, fully unrolled, registers only): triangle-inequality bounds over the
//            anchors -- FADD, FMNMX(|.|), FADD, FMNMX per anchor and pair, issue-bound -- then the
//            clipped stratified-linear prediction (dad from two shared-memory look-ups, bin by
//            compares against the doubled edges, one 128-bit coefficient fetch, 3 FFMA, clip) and
//            one compare against the pair's cut-off.  Pairs that beat the cut-off OR carry a flag
//            bit (exactly known / tightened / forced: their RefineApprox is not the plain
//            prediction) are compacted with warp ballots into a per-warp queue in shared memory;
//   phase 2 (survivors only, rolled loop, one survivor per lane): candidate test, known-pair
//            look-up, label, probability, list insertion / emission.
// In high dimension the anchor lower bound alone is a weak filter (distances concentrate: on the
// d=128 benchmark 75 % of all pairs have lb < thresh), hence the prediction is part of phase 1.
#pragma once
#include "index.cuh"

namespace annb {

// Threads per sweep CTA: 256 (each thread takes two 4 x 8 half passes of the 128 x 128 tile, <= 255
// registers) or 512 (one half pass each, <= 128 registers, 16 warps per SM to hide the fixed-latency
// and barrier stalls ncu shows at 8 warps).
#ifndef ANNB_SWEEP_THREADS
#define ANNB_SWEEP_THREADS 256
#endif
constexpr int SWT = ANNB_SWEEP_THREADS;
constexpr int SWW = SWT / 32;         // warps per sweep CTA
constexpr int SW_HALVES = 512 / SWT;  // half passes per thread
constexpr int SW_PASSES = SW_HALVES * 5;  // per half pass: 4 row steps + 1 drain-only step
constexpr int QCAP = 512;   // per-warp survivor queue; drained when a row step (<= 256 new) might overflow it
constexpr int QROW = 256;   // survivors one micro-tile row step can add: 8 pairs x 32 lanes

struct __align__(16) Survivor {
    float lb, ub;
    uint32_t ids;  // li | lj << 8
    uint32_t pad;
};

// 3-input min / max (FMNMX3 on sm_100a: two anchors per min/max instruction)
__device__ __forceinline__ float fmax3abs(float a, float b, float c)
{
    float d;
    asm("max.abs.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c)
{
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// stage the anchor-distance rows and point metadata of tile t
__device__ __forceinline__ void load_point_tile(const View &V, int t, float *sD, PointMeta *sM)
{
    const int tid = threadIdx.x;
    for (int idx = tid; idx < V.na * 32; idx += blockDim.x) {
        const int a = idx >> 5, q = idx & 31;
        cp_async16(sD + a * SROW + q * 4, V.D32 + (int64_t)a * V.npad + (int64_t)t * TILE + q * 4);
    }
    if (tid < TILE) cp_async16(sM + tid, V.meta + (int64_t)t * TILE + tid);
}
// __syncwarp() the compiler cannot elide: warp-level queues are written by one lane and read by another
// (compute-sanitizer racecheck needs to see the barrier)
__device__ __forceinline__ void warp_barrier() { asm volatile("bar.warp.sync 0xffffffff;" ::: "memory"); }
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(s), "l"(gmem) : "memory");
}

// ---- per-tile store entries -------------------------------------------------------------------
// Every pair with a store entry (exactly known / tightened / forced) is listed once, in the list of
// the upper-triangular tile that holds it: code = r << 9 | c << 2 | kind (r, c = local index of the
// lower / higher endpoint) plus the two values in parallel arrays.  A sweep stages the codes of its
// tile next to the anchor-distance rows, scatters them into a 128 x 128 flag bitmap in shared memory
// and indexes them by bit rank (rowbase + popcount), so phase 2 finds a flagged pair's entry without
// any search: no Theta(N^2) bitmap in HBM, no random hash probes in the sweeps.
constexpr int TL_CAP = 1024;  // entries staged per tile; larger tiles take their flags from the global list
                              // and their values from the hash map (small N with a large p_work)

struct TileDesc {  // [base, end) of a tile's entries in the global lists
    long long base, end;
};

// Shared-memory block of the current tile's store state: bm[BITMAP_WORDS] (flags, canonical
// orientation: row = lower endpoint) | rowbase[TILE] u16 (set bits in the rows above) | perm[TL_CAP] u16
// (bit rank -> entry) | clean flag.  The handle keeps only three pointers in registers.
constexpr int TS_BYTES = BITMAP_WORDS * 4 + TILE * 2 + TL_CAP * 2 + 16;

struct TileStore {
    uint32_t *bm;           // start of the block
    const uint32_t *codes;  // staged codes of this tile
    const TileDesc *d;      // its descriptor (shared memory)
    __device__ __forceinline__ uint16_t *rowbase() const { return reinterpret_cast<uint16_t *>(bm + BITMAP_WORDS); }
    __device__ __forceinline__ uint16_t *perm() const { return rowbase() + TILE; }
    __device__ __forceinline__ volatile uint32_t *clean() const
    {
        return reinterpret_cast<volatile uint32_t *>(perm() + TL_CAP);
    }
    __device__ __forceinline__ void init(void *block)
    {
        bm = reinterpret_cast<uint32_t *>(block);
        codes = nullptr;
        d = nullptr;
        if (threadIdx.x == 0) *clean() = 0;  // (made visible by the first barrier of the tile loop)
    }
};

// asynchronously fetch the descriptor of upper-triangular tile t (two 8-byte copies)
__device__ __forceinline__ void load_tile_desc(const View &V, int64_t t, TileDesc *d)
{
    if (threadIdx.x == 0) {
        cp_async8(&d->base, V.tl_ptr + t);
        cp_async8(&d->end, V.tl_ptr + t + 1);
    }
}
// asynchronously stage the codes of a tile whose descriptor is already in shared memory
__device__ __forceinline__ void load_tile_codes(const View &V, const TileDesc *d, uint32_t *codes)
{
    const long long base = d->base, cnt = d->end - base;
    if (cnt > TL_CAP) return;
    for (int e = threadIdx.x; e < (int)cnt; e += blockDim.x) cp_async4(codes + e, V.tl_code + base + e);
}

__device__ __forceinline__ void load_tile_codes(const View &V, long long base, long long end, uint32_t *codes)
{
    const long long cnt = end - base;
    if (cnt > TL_CAP) return;
    for (int e = threadIdx.x; e < (int)cnt; e += blockDim.x) cp_async4(codes + e, V.tl_code + base + e);
}

// `bf` is the flag bitmap phase 1 reads.  Upper-triangle sweeps pass bf == ts.bm (mode 0).  A row sweep
// keeps a separate row-oriented copy: mode 1 = transposed (row tile below the column tile), 2 = both
// orientations (diagonal tile), 3 = same orientation (row tile above the column tile).
__device__ __forceinline__ void build_tile_store(const View &V, TileStore &ts, const TileDesc *dp,
                                                 const uint32_t *codes, uint32_t *bf, int mode)
{
    const int tid = threadIdx.x;
    const long long base = dp->base;
    const int cnt = (int)(dp->end - base);
    ts.codes = codes;
    ts.d = dp;
    const bool staged = cnt <= TL_CAP;
    const bool was_clean = *ts.clean() != 0;
    if (cnt == 0 && was_clean) return;  // nothing flagged here and the bitmaps are already zero
    const bool sep = bf != ts.bm;
    for (int k = tid; k < BITMAP_WORDS; k += blockDim.x) {
        ts.bm[k] = 0;
        if (sep) bf[k] = 0;
    }
    __syncthreads();  // (also orders the read of the clean flag above before its update below)
    if (tid == 0) *ts.clean() = cnt == 0 ? 1u : 0u;
    if (cnt == 0) return;
    for (int e = tid; e < cnt; e += blockDim.x) {
        const uint32_t code = staged ? codes[e] : __ldg(V.tl_code + base + e);
        const int r = code >> 9, c = (code >> 2) & 127;
        atomicOr(&ts.bm[r * 4 + (c >> 5)], 1u << (c & 31));
        if (mode == 1 || mode == 2) atomicOr(&bf[c * 4 + (r >> 5)], 1u << (r & 31));
        if (mode == 2 || mode == 3) atomicOr(&bf[r * 4 + (c >> 5)], 1u << (c & 31));
    }
    __syncthreads();
    if (!staged) return;
    // rowbase = exclusive prefix of the per-row popcounts (threads 0..127, one row each)
    __shared__ int s_wsum[4];
    uint16_t *rowbase = ts.rowbase(), *perm = ts.perm();
    int mine = 0, incl = 0;
    if (tid < TILE) {
        const uint4 w = *reinterpret_cast<const uint4 *>(ts.bm + tid * 4);
        mine = __popc(w.x) + __popc(w.y) + __popc(w.z) + __popc(w.w);
        incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += y;
        }
        if ((tid & 31) == 31) s_wsum[tid >> 5] = incl;
    }
    __syncthreads();
    if (tid < TILE) {
        int off = 0;
        for (int q = 0; q < (tid >> 5); ++q) off += s_wsum[q];
        rowbase[tid] = (uint16_t)(off + incl - mine);
    }
    __syncthreads();
    for (int e = tid; e < cnt; e += blockDim.x) {
        const uint32_t code = codes[e];
        const int r = code >> 9, c = (code >> 2) & 127, w = c >> 5;
        const uint32_t *row = ts.bm + r * 4;
        int rho = rowbase[r] + __popc(row[w] & ((1u << (c & 31)) - 1u));
        rho += (w > 0 ? __popc(row[0]) : 0) + (w > 1 ? __popc(row[1]) : 0) + (w > 2 ? __popc(row[2]) : 0);
        perm[rho] = (uint16_t)e;
    }
    __syncthreads();
    // pull the values of this tile's entries into L2 while phase 1 runs (one line per 32 entries)
    for (int e = tid * 32; e < cnt; e += blockDim.x * 32) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(V.tl_a + base + e));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(V.tl_b + base + e));
    }
}

// store entry of a FLAGGED pair: (r, c) = local index of its lower / higher endpoint
__device__ __forceinline__ uint32_t store_lookup(const View &V, const TileStore &ts, int r, int c, uint32_t glo,
                                                 uint32_t ghi, float &a, float &b)
{
    const long long base = ts.d->base;
    if (ts.d->end - base > TL_CAP) return hash_lookup(V, pair_key(glo, ghi), a, b);
    const uint32_t *row = ts.bm + r * 4;
    const int w = c >> 5;
    int rho = ts.rowbase()[r] + __popc(row[w] & ((1u << (c & 31)) - 1u));
    rho += (w > 0 ? __popc(row[0]) : 0) + (w > 1 ? __popc(row[1]) : 0) + (w > 2 ? __popc(row[2]) : 0);
    const int e = ts.perm()[rho];
    const uint32_t kind = ts.codes[e] & 3u;
    a = __ldg(V.tl_a + base + e);
    b = kind == KIND_TIGHT ? __ldg(V.tl_b + base + e) : 0.0f;
    return kind;
}

__device__ __forceinline__ int micro_off(int t4, int k)  // local index of the k-th of 8 rows/cols
{
    return (k < 4) ? (t4 * 4 + k) : (64 + t4 * 4 + (k - 4));
}

// triangle-inequality bounds (get_bounds_njit_ijs, utils.py:274-301) for HALF of the 8 x 8
// micro-tile: 4 rows (local rows row0..row0+3) x 8 columns.  The tile is processed in two such
// passes so that the 64 accumulators, the 24 staged anchor distances and the kernel state fit the
// register file without spills.  Two anchors per step: 4 FADD + 2 FMNMX3 per pair (max / min are
// exact and order-free, so the result equals the reference's sequential loop in float32).
__device__ __forceinline__ void bounds_half(const float *__restrict__ sDi, const float *__restrict__ sDj,
                                            int na, int row0, int tx, float (&lb)[4][8], float (&ub)[4][8])
{
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            lb[r][c] = 0.0f;
            ub[r][c] = INFINITY;
        }
    int a = 0;
#ifdef ANNB_VARIANT_UNROLL2
#pragma unroll 2
#else
#pragma unroll 1
#endif
    for (; a + 2 <= na; a += 2) {
        const float *pi = sDi + a * SROW + row0, *pj = sDj + a * SROW + tx * 4;
        const float4 i0 = *reinterpret_cast<const float4 *>(pi);
        const float4 k0 = *reinterpret_cast<const float4 *>(pi + SROW);
        const float4 j0 = *reinterpret_cast<const float4 *>(pj);
        const float4 j1 = *reinterpret_cast<const float4 *>(pj + 64);
        const float4 l0 = *reinterpret_cast<const float4 *>(pj + SROW);
        const float4 l1 = *reinterpret_cast<const float4 *>(pj + SROW + 64);
        const float di[4] = {i0.x, i0.y, i0.z, i0.w};
        const float ei[4] = {k0.x, k0.y, k0.z, k0.w};
        const float dj[8] = {j0.x, j0.y, j0.z, j0.w, j1.x, j1.y, j1.z, j1.w};
        const float ej[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                lb[r][c] = fmax3abs(lb[r][c], di[r] - dj[c], ei[r] - ej[c]);
                ub[r][c] = fmin3(ub[r][c], di[r] + dj[c], ei[r] + ej[c]);
            }
    }
    if (a < na) {
        const float4 i0 = *reinterpret_cast<const float4 *>(sDi + a * SROW + row0);
        const float4 j0 = *reinterpret_cast<const float4 *>(sDj + a * SROW + tx * 4);
        const float4 j1 = *reinterpret_cast<const float4 *>(sDj + a * SROW + 64 + tx * 4);
        const float di[4] = {i0.x, i0.y, i0.z, i0.w};
        const float dj[8] = {j0.x, j0.y, j0.z, j0.w, j1.x, j1.y, j1.z, j1.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                lb[r][c] = fmaxf(lb[r][c], fabsf(di[r] - dj[c]));
                ub[r][c] = fminf(ub[r][c], di[r] + dj[c]);
            }
    }
}

// pairs of the half micro-tile that must not take part (diagonal tiles: li >= lj for the
// upper-triangle sweeps, li == lj for the row sweeps) get lb = ub = +inf: their clipped prediction is
// +inf (or NaN -> compares false), so phase 1 drops them without a per-pair test
template <bool STRICT_UPPER>
__device__ __forceinline__ void mask_diagonal(int row0, int tx, float (&lb)[4][8], float (&ub)[4][8])
{
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int li = row0 + r;
            const int lj = (c < 4) ? (tx * 4 + c) : (64 + tx * 4 + (c - 4));
            if (STRICT_UPPER ? (li >= lj) : (li == lj)) {
                lb[r][c] = INFINITY;
                ub[r][c] = INFINITY;
            }
        }
}

// shared-memory copy of the regression model in the form phase 1 consumes.  The (doubled) bin
// edges are read straight from the kernel parameters (constant bank operands: no load, no register).
struct __align__(16) TileModel {
    float4 cf[MAX_BINS];   // (c0, c1, c2 / 2, intercept) per bin (regressors.py:39-67)
    float4 mx[MAX_BINS];   // scoring only: error-floor margins of labels (b, b+1): level >= floor, level > floor
    float en[MAX_BINS];    // scoring only: doubled edge b + 1 (the pair's label is b + 1 iff 2*dad equals it)
};

__device__ __forceinline__ void build_tile_model(const Model &M, TileModel *tm)
{
    const int t = threadIdx.x;
    if (t < MAX_BINS) {
        const bool live = t < M.nb;
        tm->cf[t] = live ? make_float4(M.c0[t], M.c1[t], 0.5f * M.c2[t], M.ic[t]) : make_float4(0, 0, 0, 0);
        tm->mx[t] = make_float4(0, 0, 0, 0);
        tm->en[t] = INFINITY;
    }
}

// s = 2 * dad.  Regression bin (lo, hi]: number of interior edges strictly below dad
// (regressors.py:85-87), found with a 3-level compare tree over the 7 (padded) doubled edges M.e2.
__device__ __forceinline__ int reg_bin2(const Model &M, float s)
{
    const bool p4 = s > M.e2[4];
    const bool p2 = s > (p4 ? M.e2[6] : M.e2[2]);
    const float lo = p2 ? M.e2[3] : M.e2[1], hi = p2 ? M.e2[7] : M.e2[5];
    const bool p1 = s > (p4 ? hi : lo);
    return (p4 ? 4 : 0) + (p2 ? 2 : 0) + (p1 ? 1 : 0);
}
// error label / sampler bin: closed [lo, hi] with later bins winning (error_predictors.py:63-66):
// number of interior edges <= dad
__device__ __forceinline__ int err_label2(const Model &M, float s)
{
    const bool p4 = s >= M.e2[4];
    const bool p2 = s >= (p4 ? M.e2[6] : M.e2[2]);
    const float lo = p2 ? M.e2[3] : M.e2[1], hi = p2 ? M.e2[7] : M.e2[5];
    const bool p1 = s >= (p4 ? hi : lo);
    return (p4 ? 4 : 0) + (p2 ? 2 : 0) + (p1 ? 1 : 0);
}
// clip(lb*c0 + ub*c1 + dad*c2 + icpt, lb, ub)  (annchor.py:356-363); identical code in both phases
__device__ __forceinline__ float predict_clip2(const TileModel *tm, const Model &M, float lb, float ub,
                                               float s, int &bin)
{
    bin = reg_bin2(M, s);
    const float4 cf = tm->cf[bin];
    const float y = fmaf(lb, cf.x, fmaf(ub, cf.y, fmaf(s, cf.z, cf.w)));
    return fminf(fmaxf(y, lb), ub);
}

// the two 32-bit words of a tile bitmap row that hold this thread's 8 columns
// (columns 4*tx..4*tx+3 sit in word tx>>3, columns 64+4*tx.. in word 2 + (tx>>3), same bit offsets)
__device__ __forceinline__ void flag_words(const uint32_t *sB, int row, int tx, uint32_t &w0, uint32_t &w1)
{
    w0 = sB[row * 4 + (tx >> 3)] >> ((tx & 7) * 4);
    w1 = sB[row * 4 + 2 + (tx >> 3)] >> ((tx & 7) * 4);
}

// append one micro-tile row's survivors (bit c of km = column c survives) to the warp queue: one
// shared-memory atomic per thread that has any, then predicated 128-bit stores
__device__ __forceinline__ void stage_row(Survivor *q, int *qcnt, const float (&lb)[8], const float (&ub)[8],
                                          uint32_t km, int li, int tx)
{
    if (km == 0) return;
    int pos = atomicAdd(qcnt, __popc(km));
#pragma unroll
    for (int c = 0; c < 8; ++c)
        if ((km >> c) & 1u) {
            Survivor s;
            s.lb = lb[c];
            s.ub = ub[c];
            s.ids = (uint32_t)li | ((uint32_t)micro_off(tx, c) << 8);
            s.pad = 0;
            q[pos++] = s;
        }
}

// candidate test (get_check / adjust_check, utils.py:437-491): shared nearest-anchor count
// >= min of the two rows' thresholds
__device__ __forceinline__ bool is_candidate(const PointMeta &a, const PointMeta &b)
{
    const int t = a.loc_t < b.loc_t ? a.loc_t : b.loc_t;
    return __popcll(a.amask & b.amask) >= t;
}

// flag bit of local (row, col) in a tile bitmap
__device__ __forceinline__ bool flag_bit(const uint32_t *sB, int r, int c)
{
    return (sB[r * 4 + (c >> 5)] >> (c & 31)) & 1u;
}

// (ti, tj), ti <= tj, of the t-th upper-triangular tile (inverse of tile_index)
__device__ __forceinline__ void tile_from_index(int64_t t, int T, int &ti, int &tj)
{
    // invert tile_index(): largest ti with ti*T - ti*(ti-1)/2 <= t
    const double b = 2.0 * T + 1.0;
    int r = (int)((b - sqrt(b * b - 8.0 * (double)t)) * 0.5);
    if (r < 0) r = 0;
    if (r > T - 1) r = T - 1;
    while (r > 0 && tile_index(r, r, T) > t) --r;
    while (r + 1 < T && tile_index(r + 1, r + 1, T) <= t) ++r;
    ti = r;
    tj = r + (int)(t - tile_index(r, r, T));
}

// ---- tile-level pruning ------------------------------------------------------------------------
// Lower bound of the clipped prediction over ALL pairs (i in tile ti, j in tile tj), from the per-tile
// min / max anchor distances: interval arithmetic on the very operations phase 1 performs
//   lb = max_a |D_ia - D_ja| >= max_a gap_a,   ub = min_a (D_ia + D_ja) >= min_a (lo_a + lo'_a),
//   s2 = D_i,cA(j) + D_j,cA(i)  in  [s_lo, s_hi]   (over the closest anchors that occur in the tiles),
//   y  = fma(lb, c0, fma(ub, c1, fma(s2, c2/2, icpt)))  evaluated at the interval end the sign of each
//        coefficient selects, minimised over the regression bins the s2 interval overlaps,
// then clipped.  Floating-point subtraction, addition, fma, min and max are monotone, so the result is
// a true lower bound of what predict_clip2 returns for any pair of the two tiles -- no tolerance needed.
// Every warp computes it redundantly (lanes over anchors, ~150 instructions per tile).
// Can any pair of tiles (ti, tj) pass the phase-1 test?  MODE 0 (scoring): cut - pred > margin[label];
// MODE 1 (pair thresholds): pred <= cut; MODE 2 (row thresholds): pred < cut -- with cut <= cutmax and
// pred >= the bound below.  The bound and the margin are taken per regression bin the tile pair's s2
// interval overlaps (the error label of a pair is its bin b, or b + 1 when s2 sits exactly on the edge).
// per regression bin the interval [s_lo, s_hi] of 2 * dad overlaps: lower bound of the clipped prediction from the
// interval ends the coefficient signs select, compared with the cut the way phase 1 compares a pair
template <int MODE>
__device__ __forceinline__ bool bins_can_pass(const Model &M, float lbmin, float lbmax, float ubmin, float ubmax,
                                              float s_lo, float s_hi, float cutmax, const float *margin)
{
    bool any = false;
    for (int b = 0; b < M.nb; ++b) {
        // bin b holds e2[b] < s2 <= e2[b+1]  (e2[0] = -inf, e2[nb] = +inf; reg_bin2)
        const float blo = b == 0 ? -INFINITY : M.e2[b], bhi = b + 1 < M.nb ? M.e2[b + 1] : INFINITY;
        if (!(s_hi > blo && s_lo <= bhi)) continue;
        const float c0 = M.c0[b], c1 = M.c1[b], cz = 0.5f * M.c2[b];
        const float y = fmaf(c0 >= 0.0f ? lbmin : lbmax, c0,
                             fmaf(c1 >= 0.0f ? ubmin : ubmax, c1, fmaf(cz >= 0.0f ? s_lo : s_hi, cz, M.ic[b])));
        const float p = fminf(fmaxf(y, lbmin), ubmin);  // lower bound of the clipped prediction in this bin
        if (MODE == 0) {
            const float mg = fminf(margin[b], margin[b + 1 < M.nb ? b + 1 : b]);
            any |= !(cutmax - p <= mg);  // (NaN counts as "can pass")
        } else if (MODE == 1) {
            any |= !(p > cutmax);
        } else {
            any |= !(p >= cutmax);
        }
    }
    return any;
}

template <int MODE>
__device__ __forceinline__ bool tile_can_pass(const View &V, const Model &M, int ti, int tj, float cutmax,
                                              const float *margin /* [MAX_BINS], MODE 0 */)
{
    const int lane = threadIdx.x & 31;
    const float *loI = V.tb_lo + (int64_t)ti * kMaxAnchors, *hiI = V.tb_hi + (int64_t)ti * kMaxAnchors;
    const float *loJ = V.tb_lo + (int64_t)tj * kMaxAnchors, *hiJ = V.tb_hi + (int64_t)tj * kMaxAnchors;
    const uint64_t cmI = V.tb_cm[ti], cmJ = V.tb_cm[tj];
    float lbmin = 0.0f, lbmax = 0.0f, ubmin = INFINITY, ubmax = INFINITY;
    float siLo = INFINITY, siHi = -INFINITY, sjLo = INFINITY, sjHi = -INFINITY;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int a = lane + 32 * q;
        if (a < V.na) {
            const float li = loI[a], hi = hiI[a], lj = loJ[a], hj = hiJ[a];
            lbmin = fmaxf(lbmin, fmaxf(li - hj, lj - hi));
            lbmax = fmaxf(lbmax, fmaxf(hi - lj, hj - li));
            ubmin = fminf(ubmin, li + lj);
            ubmax = fminf(ubmax, hi + hj);
            if ((cmJ >> a) & 1ull) {  // D_i,cA(j): anchors that are the closest one of some point of tile j
                siLo = fminf(siLo, li);
                siHi = fmaxf(siHi, hi);
            }
            if ((cmI >> a) & 1ull) {
                sjLo = fminf(sjLo, lj);
                sjHi = fmaxf(sjHi, hj);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lbmin = fmaxf(lbmin, __shfl_xor_sync(0xffffffffu, lbmin, o));
        lbmax = fmaxf(lbmax, __shfl_xor_sync(0xffffffffu, lbmax, o));
        ubmin = fminf(ubmin, __shfl_xor_sync(0xffffffffu, ubmin, o));
        ubmax = fminf(ubmax, __shfl_xor_sync(0xffffffffu, ubmax, o));
        siLo = fminf(siLo, __shfl_xor_sync(0xffffffffu, siLo, o));
        siHi = fmaxf(siHi, __shfl_xor_sync(0xffffffffu, siHi, o));
        sjLo = fminf(sjLo, __shfl_xor_sync(0xffffffffu, sjLo, o));
        sjHi = fmaxf(sjHi, __shfl_xor_sync(0xffffffffu, sjHi, o));
    }
    return bins_can_pass<MODE>(M, lbmin, lbmax, ubmin, ubmax, siLo + sjLo, siHi + sjHi, cutmax, margin);
}

// The same question for ONE point (local row l of a staged tile, anchor distances in shared memory) against
// a whole tile t: can any pair (point, member of t) pass phase 1 through a cut <= `cut`?  Per-thread, no
// shuffles; the bounds of the point's own coordinates are exact, so this is much sharper than the tile-tile test.
template <int MODE>
__device__ __forceinline__ bool point_can_pass(const View &V, const Model &M, const float *__restrict__ sD, int l,
                                               int cA_l, int t, float cut, const float *margin)
{
    const float *lo = V.tb_lo + (int64_t)t * kMaxAnchors, *hi = V.tb_hi + (int64_t)t * kMaxAnchors;
    const uint64_t cm = V.tb_cm[t];
    float lbmin = 0.0f, lbmax = 0.0f, ubmin = INFINITY, ubmax = INFINITY, siLo = INFINITY, siHi = -INFINITY;
#pragma unroll 6
    for (int a = 0; a < V.na; ++a) {
        const float d = sD[a * SROW + l], lj = __ldg(lo + a), hj = __ldg(hi + a);
        lbmin = fmaxf(lbmin, fmaxf(d - hj, lj - d));
        lbmax = fmaxf(lbmax, fmaxf(hj - d, d - lj));
        ubmin = fminf(ubmin, d + lj);
        ubmax = fminf(ubmax, d + hj);
        if ((cm >> a) & 1ull) {  // D[point, cA(j)] for the closest anchors that occur in t
            siLo = fminf(siLo, d);
            siHi = fmaxf(siHi, d);
        }
    }
    // D[j, cA(point)]: the tile's interval on the point's own closest anchor
    const float sjLo = __ldg(lo + cA_l), sjHi = __ldg(hi + cA_l);
    return bins_can_pass<MODE>(M, lbmin, lbmax, ubmin, ubmax, siLo + sjLo, siHi + sjHi, cut, margin);
}

// ---- reduced tile mode -------------------------------------------------------------------------
// A tile's cut-offs are not uniform: a few rows (points whose neighbourhood the pre-pass missed, rows
// without enough candidates: cut = +inf) sit far above the rest, and the tile-level test against the LARGEST
// cut of 128 rows then keeps alive a tile pair that only those rows can use.  So a tile pair that survives
// the test is examined per row: tile_can_pass with the row's OWN cut tells whether any pair of that row can
// pass phase 1 through the row's cut (a pair passes through max(cut_i, cut_j), i.e. through one of its two
// rows).  When at most MAXO rows of either tile can, the sweep computes only (those rows x all columns),
// (all rows x those columns) and the pairs with a store entry -- a few hundred pairs instead of 16 384 --
// with the same per-pair arithmetic and the same phase 2.
constexpr int MAXO = 16;
constexpr int REDUCED_MAX_ITEMS = 3072;  // pairs of a reduced tile (outlier rows x 128 + store entries) above which the full tile is cheaper
struct Outliers {  // shared memory
    uint32_t mi[4], mj[4];   // rows of the row tile / the column tile that can pass, as bit masks
    int n_i, n_j;
    uint8_t oi[MAXO], oj[MAXO];
};
// Every thread calls this with `mine` = "my row can pass" (threads 0..127: rows of the row tile, 128..255: rows of the
// column tile).  Returns true, with the index lists filled, if both tiles have at most MAXO such rows.  Block-wide.
__device__ __forceinline__ bool collect_outliers(Outliers *o, bool mine)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned bal = __ballot_sync(0xffffffffu, mine);
    if (lane == 0 && warp < 8) (warp < 4 ? o->mi[warp] : o->mj[warp - 4]) = bal;
    __syncthreads();
    const int n_i = __popc(o->mi[0]) + __popc(o->mi[1]) + __popc(o->mi[2]) + __popc(o->mi[3]);
    const int n_j = __popc(o->mj[0]) + __popc(o->mj[1]) + __popc(o->mj[2]) + __popc(o->mj[3]);
    const bool ok = n_i <= MAXO && n_j <= MAXO;
    if (ok && warp < 2) {  // warp 0 lists the rows, warp 1 the columns: lane l takes the l-th set bit
        const uint32_t *m = warp == 0 ? o->mi : o->mj;
        int k = lane, idx = -1;
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int c = __popc(m[w]);
            if (idx < 0) {
                if (k < c) idx = w * 32 + (int)__fns(m[w], 0, k + 1);
                else k -= c;
            }
        }
        if (lane < (warp == 0 ? n_i : n_j)) (warp == 0 ? o->oi : o->oj)[lane] = (uint8_t)idx;
        if (lane == 0) (warp == 0 ? o->n_i : o->n_j) = warp == 0 ? n_i : n_j;
    }
    __syncthreads();
    return ok;
}
// no row can pass on its own: only the store entries of the tile are left
__device__ __forceinline__ void no_outliers(Outliers *o)
{
    if (threadIdx.x < 4) {
        o->mi[threadIdx.x] = 0u;
        o->mj[threadIdx.x] = 0u;
    }
    if (threadIdx.x == 0) o->n_i = o->n_j = 0;
    __syncthreads();
}
// bounds of ONE pair from the staged anchor-distance tiles: one rounded subtraction / addition per anchor and
// exact max / min, i.e. bit-identical to what bounds_half leaves for the same pair
__device__ __forceinline__ void bounds_pair(const float *__restrict__ sDi, const float *__restrict__ sDj, int na, int li,
                                            int lj, float &lb, float &ub)
{
    lb = 0.0f;
    ub = INFINITY;
    for (int a = 0; a < na; ++a) {
        const float x = sDi[a * SROW + li], y = sDj[a * SROW + lj];
        lb = fmaxf(lb, fabsf(x - y));
        ub = fminf(ub, x + y);
    }
}
// item `idx` of a reduced tile: [0, 128 n_i) outlier rows x columns | [.., + 128 n_j) rows x outlier columns (rows
// that are outliers themselves are covered by the first group) | entries of the store (not in an outlier row /
// column).  Returns false for an index that is not a pair of its own.
__device__ __forceinline__ bool reduced_item(const View &V, const Outliers *o, const TileDesc *dp, const uint32_t *codes,
                                             int idx, int &li, int &lj)
{
    const int na_ = o->n_i * TILE, nb_ = o->n_j * TILE;
    if (idx < na_) {
        li = o->oi[idx >> 7];
        lj = idx & 127;
        return true;
    }
    if (idx < na_ + nb_) {
        const int k = idx - na_;
        lj = o->oj[k >> 7];
        li = k & 127;
        return ((o->mi[li >> 5] >> (li & 31)) & 1u) == 0u;
    }
    const long long e = idx - na_ - nb_;
    const long long cnt = dp->end - dp->base;
    if (e >= cnt) return false;
    const uint32_t code = cnt <= TL_CAP ? codes[e] : __ldg(V.tl_code + dp->base + e);
    li = code >> 9;
    lj = (code >> 2) & 127;
    return ((o->mi[li >> 5] >> (li & 31)) & 1u) == 0u && ((o->mj[lj >> 5] >> (lj & 31)) & 1u) == 0u;
}

// largest of 128 shared-memory values (every warp computes it)
__device__ __forceinline__ float tile_max128(const float *v)
{
    const int lane = threadIdx.x & 31;
    float m = fmaxf(fmaxf(v[lane], v[lane + 32]), fmaxf(v[lane + 64], v[lane + 96]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    return m;
}

struct PairVal {
    float v;        // RefineApprox value (exact if computed, clipped prediction otherwise, -1 if forced)
    float dad;      // 2 * double anchor distance
    bool computed;  // anchor pair or exactly evaluated (not_computed_mask == False)
};

// RefineApprox of one candidate pair as annchor.py:345-380 leaves it.  (cr, cc) = local index of the
// pair's lower / higher endpoint inside the canonical (upper-triangular) tile, used when flagged.
__device__ __forceinline__ PairVal pair_value(const View &V, const TileStore &ts, const TileModel *tm, const Model &M,
                                              float lb, float ub, int li, int lj, int gi, int gj, int cr, int cc,
                                              const PointMeta &pi, const PointMeta &pj, const float *sDi,
                                              const float *sDj, bool flagged)
{
    PairVal out;
    const float s = sDi[pj.cA * SROW + li] + sDj[pi.cA * SROW + lj];  // 2 * dad (utils.py:378-380)
    out.dad = s;
    const bool anchorpair = (pi.slot >= 0) | (pj.slot >= 0);
    out.computed = anchorpair;
    if (flagged) {
        float a = 0.0f, b = 0.0f;
        const uint32_t lo = gi < gj ? gi : gj, hi = gi < gj ? gj : gi;
        const uint32_t kind = store_lookup(V, ts, cr, cc, lo, hi, a, b);
        if (kind == KIND_KNOWN) {
            out.v = a;
            out.computed = true;
            return out;
        }
        if (kind == KIND_TIGHT) {  // update_anchor_points, annchor.py:503-510
            lb = fmaxf(lb, a);
            ub = fminf(ub, b);
        } else if (kind == KIND_FORCED) {  // guarantee_nmin, utils.py:619
            out.v = -1.0f;
            return out;
        }
    }
    int bin;
    out.v = predict_clip2(tm, M, lb, ub, s, bin);
    if (!V.is_metric && anchorpair) {
        // annchor.py:368-372: anchor distances written explicitly; the later anchor in A wins
        out.v = (pj.slot > pi.slot) ? sDi[pj.slot * SROW + li] : sDj[pi.slot * SROW + lj];
    }
    return out;
}

}  // namespace annb
