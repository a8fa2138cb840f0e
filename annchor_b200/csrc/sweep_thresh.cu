// sweep_thresh.cu -- K5a: per-row selection thresholds without materialising RefineApprox.
//
//   thresh[i] = (nn+1)-th smallest RefineApprox over the candidate pairs of row i
//               (annchor/annchor.py:399-404)
// and, on iteration 0, the inputs of guarantee_nmin (annchor/utils.py:606-621): per row the
// nmin+1 smallest NOT-computed predictions with their column ids.
//
// One CTA owns a block of 128 rows and streams every column tile past it (both triangles:
// each pair is evaluated once per endpoint, which keeps all per-row state private -- a warp
// owns 16 whole rows of the tile, so the sorted row lists in shared memory need no inter-warp
// synchronisation, no global atomics, and the result is deterministic).  A pair is handed to
// phase 2 only when its lower bound beats the row's current k-th value, so after the first few
// column tiles almost nothing survives phase 1.
#include <algorithm>

#include "sweep.cuh"
#include "sweep_args.cuh"

namespace annb {


struct SortedRow {
    // warp-cooperative sorted insert into a list of K <= 64 (value, id) pairs kept in shared memory
    __device__ static __forceinline__ void insert(float *lv, int32_t *li, int K, float v, int32_t id,
                                                  int lane, float *thr)
    {
        const bool h0 = lane < K, h1 = lane + 32 < K;
        const float x0 = h0 ? lv[lane] : INFINITY, x1 = h1 ? lv[lane + 32] : INFINITY;
        int32_t a0 = 0, a1 = 0;
        if (li) {
            a0 = h0 ? li[lane] : INT32_MAX;
            a1 = h1 ? li[lane + 32] : INT32_MAX;
        }
        const bool le0 = (x0 < v) || (x0 == v && (!li || a0 <= id));
        const bool le1 = (x1 < v) || (x1 == v && (!li || a1 <= id));
        const int pos = __popc(__ballot_sync(0xffffffffu, le0 && h0)) +
                        __popc(__ballot_sync(0xffffffffu, le1 && h1));
        if (pos >= K) return;
        const float y0 = __shfl_up_sync(0xffffffffu, x0, 1), y1 = __shfl_up_sync(0xffffffffu, x1, 1);
        const float carry = __shfl_sync(0xffffffffu, x0, 31);
        const float n0 = lane < pos ? x0 : (lane == pos ? v : y0);
        const int i1 = lane + 32;
        const float n1 = i1 < pos ? x1 : (i1 == pos ? v : (lane == 0 ? carry : y1));
        if (li) {
            const int32_t b0 = __shfl_up_sync(0xffffffffu, a0, 1), b1 = __shfl_up_sync(0xffffffffu, a1, 1);
            const int32_t ca = __shfl_sync(0xffffffffu, a0, 31);
            const int32_t m0 = lane < pos ? a0 : (lane == pos ? id : b0);
            const int32_t m1 = i1 < pos ? a1 : (i1 == pos ? id : (lane == 0 ? ca : b1));
            if (h0) li[lane] = m0;
            if (h1) li[i1] = m1;
        }
        if (h0) lv[lane] = n0;
        if (h1) lv[i1] = n1;
        if (lane == ((K - 1) & 31)) *thr = (K - 1 < 32) ? n0 : n1;
        __syncwarp();
    }
};

template <bool L2ON>
__global__ void __launch_bounds__(SWT, 1) thresh_sweep_kernel(const __grid_constant__ ThreshArgs A)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const View &V = A.V;
    const Model &M = A.M;
    const int na = V.na;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    // ---- shared memory carve-up ----
    float *sDi = reinterpret_cast<float *>(smem);
    float *sDj0 = sDi + na * SROW;
    float *sDj1 = sDj0 + na * SROW;
    PointMeta *sMi = reinterpret_cast<PointMeta *>(sDj1 + na * SROW);
    PointMeta *sMj0 = sMi + TILE;
    PointMeta *sMj1 = sMj0 + TILE;
    unsigned char *sTS = reinterpret_cast<unsigned char *>(sMj1 + TILE);  // store state (flags in canonical orientation)
    uint32_t *sBF = reinterpret_cast<uint32_t *>(sTS + TS_BYTES);  // flags oriented (row = this CTA's point, col = column point)
    uint32_t *sCode = sBF + BITMAP_WORDS;                          // [2][TL_CAP] staged store codes
    TileDesc *sDesc = reinterpret_cast<TileDesc *>(sCode + 2 * TL_CAP);  // [4]
    TileModel *tm = reinterpret_cast<TileModel *>(sDesc + 4);
    Survivor *queue = reinterpret_cast<Survivor *>(tm + 1) + warp * A.qcap;
    float *thr1 = reinterpret_cast<float *>(reinterpret_cast<Survivor *>(tm + 1) + SWW * A.qcap);
    float *thr2 = thr1 + TILE;
    float *cut = thr2 + TILE;  // max(thr1, thr2): phase-1 cut-off per row
    int *qcnt = reinterpret_cast<int *>(cut + TILE) + warp;  // [SWW]
    float *L1 = reinterpret_cast<float *>(reinterpret_cast<int *>(cut + TILE) + SWW);
    float *L2v = L1 + TILE * A.k1;
    int32_t *L2i = reinterpret_cast<int32_t *>(L2v + TILE * A.k2);
    const bool filter = V.is_metric != 0;  // is_metric=False overrides anchor pairs with raw D (annchor.py:368-372)

    build_tile_model(M, tm);
    if (lane == 0) *qcnt = 0;
    TileStore ts;
    ts.init(sTS);
    __shared__ unsigned char s_flag[2][SWW];
    int scan_round = 0;
    const int n_rb = A.rb_list ? A.n_rb : V.T;
    for (int rq = blockIdx.x * A.world + A.rank; rq < n_rb; rq += gridDim.x * A.world) {
        const int rb = A.rb_list ? A.rb_list[rq] : rq;
        // ---- init row state, stage the row tile ----
        for (int k = tid; k < TILE * A.k1; k += blockDim.x) L1[k] = INFINITY;
        if (L2ON)
            for (int k = tid; k < TILE * A.k2; k += blockDim.x) {
                L2v[k] = INFINITY;
                L2i[k] = -1;
            }
        if (tid < TILE) {
            thr1[tid] = INFINITY;
            thr2[tid] = L2ON ? INFINITY : -INFINITY;
            cut[tid] = INFINITY;
        }
        // canonical (upper-triangular) tile that holds the pairs of (row block rb, column tile tc)
        auto canon = [&](int tc) { return tile_index(rb < tc ? rb : tc, rb < tc ? tc : rb, V.T); };
        // column tile sequence: the band around the diagonal first (if any), then col_phase + k * col_stride
        // outside the band; -1 past the end
        const int blo = A.band > 0 ? max(0, rb - A.band) : 0, bhi = A.band > 0 ? min(V.T - 1, rb + A.band) : -1;
        const int nband = bhi - blo + 1;
        const int kb = (nband == 0 || blo <= A.col_phase) ? 0 : (blo - A.col_phase + A.col_stride - 1) / A.col_stride;
        const int ka = (nband == 0 || bhi < A.col_phase) ? 0 : (bhi - A.col_phase) / A.col_stride + 1;
        auto col_at = [&](int s) {
            int tc;
            if (s < nband) tc = blo + s;
            else {
                const int e = s - nband;
                tc = A.col_phase + ((nband == 0 || e < kb) ? e : ka + (e - kb)) * A.col_stride;
            }
            return tc < V.T ? tc : -1;
        };
        __syncthreads();  // the previous row block is done with the descriptors and the cuts are initialised
        // scan-ahead over the column tile sequence (see thresh_pairs_kernel): a column tile without store entries that
        // fails the tile test against the largest CURRENT cut of the 128 rows is skipped before it is loaded (cuts
        // only ever decrease, so a test made ahead of time is conservative)
        const bool cullable = V.cull && filter;
        auto next_step = [&](int start) -> int {
            if (!cullable) return start;
            for (int base = start;; base += SWW, ++scan_round) {
                if (col_at(base) < 0) return base;
                const int tc = col_at(base + warp);
                bool surv = false;
                if (tc >= 0) {
                    const int64_t t = canon(tc);
                    surv = tc == rb || V.tl_ptr[t + 1] > V.tl_ptr[t] ||
                           tile_can_pass<2>(V, M, rb, tc, tile_max128(cut), nullptr);
                }
                if (lane == 0) s_flag[scan_round & 1][warp] = surv ? 1 : 0;
                __syncthreads();
                int first = -1;
#pragma unroll
                for (int w = SWW - 1; w >= 0; --w)
                    if (s_flag[scan_round & 1][w]) first = w;
                if (first >= 0) {
                    ++scan_round;
                    return base + first;
                }
            }
        };
        int step_no = next_step(0), step_next = col_at(step_no) >= 0 ? next_step(step_no + 1) : step_no, step_nn = 0;
        load_point_tile(V, rb, sDi, sMi);
        if (col_at(step_no) >= 0) {
            const int64_t t = canon(col_at(step_no));
            const long long e0 = V.tl_ptr[t], e1 = V.tl_ptr[t + 1];
            if (tid == 0) {
                sDesc[0].base = e0;
                sDesc[0].end = e1;
            }
            load_point_tile(V, col_at(step_no), sDj0, sMj0);
            load_tile_codes(V, e0, e1, sCode);
        }
        cp_async_commit();

        for (int k = 0; col_at(step_no) >= 0; step_no = step_next, step_next = step_nn, ++k) {
            const int par = k & 1;  // buffer parity
            const int tc = col_at(step_no);
            const float *sDj = par ? sDj1 : sDj0;
            const PointMeta *sMj = par ? sMj1 : sMj0;
            cp_async_wait_all();
            __syncthreads();
            if (col_at(step_next) >= 0) {  // prefetch the next surviving column tile into the other buffer
                const int64_t t = canon(col_at(step_next));
                const long long e0 = V.tl_ptr[t], e1 = V.tl_ptr[t + 1];
                if (tid == 0) {
                    sDesc[(k + 1) & 3].base = e0;
                    sDesc[(k + 1) & 3].end = e1;
                }
                load_point_tile(V, col_at(step_next), par ? sDj0 : sDj1, par ? sMj0 : sMj1);
                load_tile_codes(V, e0, e1, sCode + (par ^ 1) * TL_CAP);
                cp_async_commit();
            }
            step_nn = col_at(step_next) >= 0 ? next_step(step_next + 1) : step_next;
            // entries are stored once per pair, in the tile of (lo, hi): the flags phase 1 reads are
            // re-oriented so that bit (row, col) of sBF is the flag of (this CTA's row point, column point)
            // tile-level pruning (metrics only: the phase-1 filter `pred < cut[row]` is on): no store entry
            // and the smallest possible prediction is not below the largest current cut of the 128 rows
            if (V.cull && filter && sDesc[k & 3].end == sDesc[k & 3].base && tc != rb) {
                if (!tile_can_pass<2>(V, M, rb, tc, tile_max128(cut), nullptr)) continue;
            }
            build_tile_store(V, ts, &sDesc[k & 3], sCode + par * TL_CAP, sBF, tc < rb ? 1 : (tc == rb ? 2 : 3));
            // ---- phase 1: bounds + clipped prediction, two passes of 4 x 8 pairs per thread ----
            int cAj[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) cAj[c] = sMj[micro_off(tx, c)].cA * SROW;
            float lb[4][8], ub[4][8];
#pragma unroll 1
            for (int pass = 0; pass < SW_PASSES; ++pass) {  // per half pass: 4 row steps + 1 drain-only step
                const int h = pass >= 5 ? 1 : 0, step = pass - 5 * h;
                const int row0 = h * (SWT / 4) + ty * 4;
                if (step == 0) {
                    bounds_half(sDi, sDj, na, row0, tx, lb, ub);
                    if (tc == rb) mask_diagonal<false>(row0, tx, lb, ub);
                }
                auto row_step = [&](const float (&lbr)[8], const float (&ubr)[8], int r) {
                    const int li = row0 + r;
                    const float cr = filter ? cut[li] : INFINITY;
                    const float *dj_row = sDj + sMi[li].cA * SROW;
                    uint32_t w0, w1;
                    flag_words(sBF, li, tx, w0, w1);
                    uint32_t km = (w0 & 0xfu) | ((w1 & 0xfu) << 4);  // flagged pairs always go to phase 2
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float s2 = sDi[cAj[c] + li] + dj_row[micro_off(tx, c)];
                        int bin;
                        const float y = predict_clip2(tm, M, lbr[c], ubr[c], s2, bin);
                        km |= (y < cr) ? (1u << c) : 0u;
                    }
                    stage_row(queue, qcnt, lbr, ubr, km, li, tx);
                };
                switch (step) {
                    case 0: row_step(lb[0], ub[0], 0); break;
                    case 1: row_step(lb[1], ub[1], 1); break;
                    case 2: row_step(lb[2], ub[2], 2); break;
                    case 3: row_step(lb[3], ub[3], 3); break;
                    default: break;
                }
                __syncwarp();
                const int qn = *reinterpret_cast<volatile int *>(qcnt);
                if (pass < SW_PASSES - 1 && qn <= A.qcap - QROW) continue;  // room for another row step
                // ---- phase 2: drain the queue, one survivor per lane ----
                for (int e0 = 0; e0 < qn; e0 += 32) {
                    const int e = e0 + lane;
                    bool w1b = false, w2b = false;
                    float v = INFINITY;
                    int row = 0, gj = 0;
                    if (e < qn) {
                        const Survivor sv = queue[e];
                        row = sv.ids & 0xff;
                        const int lj = sv.ids >> 8;
                        const int gi = rb * TILE + row;
                        gj = tc * TILE + lj;
                        const PointMeta pi = sMi[row], pj = sMj[lj];
                        if (gi < V.n && gj < V.n && gi != gj && is_candidate(pi, pj)) {
                            const bool fl = flag_bit(sBF, row, lj);
                            const bool rlo = gi < gj;  // canonical local coordinates: lower endpoint first
                            const PairVal pv = pair_value(V, ts, tm, M, sv.lb, sv.ub, row, lj, gi, gj, rlo ? row : lj,
                                                          rlo ? lj : row, pi, pj, sDi, sDj, fl);
                            v = pv.v;
                            w1b = v < thr1[row];
                            w2b = L2ON && !pv.computed && v < thr2[row];
                        }
                    }
                    unsigned m1 = __ballot_sync(0xffffffffu, w1b);
                    __syncwarp();  // every lane has read thr1 / thr2 before an insert rewrites them
                    while (m1) {
                        const int src = __ffs(m1) - 1;
                        m1 &= m1 - 1;
                        const float xv = __shfl_sync(0xffffffffu, v, src);
                        const int sr = __shfl_sync(0xffffffffu, row, src);
                        SortedRow::insert(L1 + sr * A.k1, nullptr, A.k1, xv, 0, lane, &thr1[sr]);
                    }
                    if (L2ON) {
                        unsigned m2 = __ballot_sync(0xffffffffu, w2b);
                        while (m2) {
                            const int src = __ffs(m2) - 1;
                            m2 &= m2 - 1;
                            const float xv = __shfl_sync(0xffffffffu, v, src);
                            const int sr = __shfl_sync(0xffffffffu, row, src);
                            const int sj = __shfl_sync(0xffffffffu, gj, src);
                            SortedRow::insert(L2v + sr * A.k2, L2i + sr * A.k2, A.k2, xv, sj, lane, &thr2[sr]);
                        }
                    }
                }
                __syncwarp();
                // refresh the phase-1 cut-offs of this warp's 16 rows ({8w..8w+7} U {64+8w..})
                if (lane < 8 * SW_HALVES) {  // this warp's rows: {8w..8w+7} (and {64+8w..} with 256 threads)
                    const int row = (lane < 8) ? (warp * 8 + lane) : (64 + warp * 8 + (lane - 8));
                    cut[row] = fmaxf(thr1[row], thr2[row]);
                }
                if (lane == 0) *qcnt = 0;
                __syncwarp();
            }
        }

        // ---- write the row block's results ----
        __syncthreads();
        if (tid < TILE) A.thresh[rb * TILE + tid] = L1[tid * A.k1 + (A.k1 - 1)];
        if (L2ON && A.cut2 && tid < TILE) A.cut2[rb * TILE + tid] = L2v[tid * A.k2 + (A.k2 - 1)];
        if (L2ON) {
            for (int k = tid; k < TILE * A.k2; k += blockDim.x) {
                const int gi = rb * TILE + k / A.k2;
                if (gi < V.n) {
                    A.l2val[(int64_t)gi * A.k2 + k % A.k2] = L2v[k];
                    A.l2id[(int64_t)gi * A.k2 + k % A.k2] = L2i[k];
                }
            }
        }
        __syncthreads();
    }
}

static size_t thresh_smem_base(int na, int k1, int k2)
{
    return (size_t)3 * na * SROW * 4 + 3 * TILE * sizeof(PointMeta) + TS_BYTES + BITMAP_WORDS * 4 + 2 * TL_CAP * 4 +
           4 * sizeof(TileDesc) + sizeof(TileModel) + 3 * TILE * 4 + SWW * 4 + (size_t)TILE * k1 * 4 + (size_t)TILE * k2 * 8 + 64;
}

int launch_thresh_sweep(annb_ctx *c, ThreshArgs &A)
{
    const size_t lim = 227 * 1024;
    const size_t base = thresh_smem_base(A.V.na, A.k1, A.k2);
    ANNB_REQUIRE(base + (size_t)SWW * (QROW + 32) * sizeof(Survivor) <= lim, ANNB_ERANGE,
                 "thresh sweep needs %zu bytes of shared memory (n_anchors=%d, lists %d/%d)",
                 base + (size_t)SWW * (QROW + 32) * sizeof(Survivor), A.V.na, A.k1, A.k2);
    ANNB_REQUIRE(A.k1 <= MAX_LIST && A.k2 <= MAX_LIST, ANNB_ERANGE,
                 "n_neighbors too large for the device row lists (k1=%d, k2=%d, max %d)", A.k1, A.k2,
                 MAX_LIST);
    int qcap = (int)((lim - base) / (SWW * sizeof(Survivor))) / 32 * 32;
    if (qcap > QCAP) qcap = QCAP;
    A.qcap = qcap;
    const size_t smem = base + (size_t)SWW * qcap * sizeof(Survivor);
    const int n_rb = A.rb_list ? A.n_rb : A.V.T;
    const int rows_here = (n_rb - A.rank + A.world - 1) / A.world;
    const int grid = rows_here < 1 ? 1 : rows_here;
    if (A.col_stride < 1) A.col_stride = 1;
    if (A.k2 > 0) {
        ANNB_CUDA(cudaFuncSetAttribute(thresh_sweep_kernel<true>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ANNB_LAUNCH(thresh_sweep_kernel<true>, grid, SWT, smem, c->stream, A);
    } else {
        ANNB_CUDA(cudaFuncSetAttribute(thresh_sweep_kernel<false>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ANNB_LAUNCH(thresh_sweep_kernel<false>, grid, SWT, smem, c->stream, A);
    }
    return ANNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Two-stage thresholds.  The row sweep above evaluates every pair twice (once per endpoint).  For
// metrics, a pre-pass of it over every S-th column tile gives, per point, the k-th smallest value of
// a SUBSET of its pairs -- an upper bound of its threshold.  The pass below then visits each pair
// ONCE (upper-triangular tiles, like the scoring sweep) and appends (value, other endpoint) to the
// record list of either endpoint whose bound it does not exceed; thresh_select_kernel finishes per
// row.  Rows whose record list overflowed are recomputed by the row sweep (launch with rb_list).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SWT, 1) thresh_pairs_kernel(const __grid_constant__ ThreshPairArgs A)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const View &V = A.V;
    const Model &M = A.M;
    const int na = V.na;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    float *sD0i = reinterpret_cast<float *>(smem);
    float *sD0j = sD0i + na * SROW;
    float *sD1i = sD0j + na * SROW;
    float *sD1j = sD1i + na * SROW;
    PointMeta *sM0i = reinterpret_cast<PointMeta *>(sD1j + na * SROW);
    PointMeta *sM0j = sM0i + TILE;
    PointMeta *sM1i = sM0j + TILE;
    PointMeta *sM1j = sM1i + TILE;
    unsigned char *sTS = reinterpret_cast<unsigned char *>(sM1j + TILE);
    uint32_t *sBm = reinterpret_cast<uint32_t *>(sTS);
    uint32_t *sCode = reinterpret_cast<uint32_t *>(sTS + TS_BYTES);  // [2][TL_CAP]
    TileDesc *sDesc = reinterpret_cast<TileDesc *>(sCode + 2 * TL_CAP);  // [4]
    Survivor *queue = reinterpret_cast<Survivor *>(sDesc + 4) + warp * A.qcap;
    float *c1I = reinterpret_cast<float *>(reinterpret_cast<Survivor *>(sDesc + 4) + SWW * A.qcap);
    float *c1J = c1I + 2 * TILE;  // [2][128] each: cut1 / cut2 of the row and column tiles (double-buffered)
    float *c2I = c1J + 2 * TILE;
    float *c2J = c2I + 2 * TILE;
    TileModel *tm = reinterpret_cast<TileModel *>(c2J + 2 * TILE);
    int *qcnt = reinterpret_cast<int *>(tm + 1) + warp;
    const bool l2on = A.cut2 != nullptr;
    build_tile_model(M, tm);
    if (lane == 0) *qcnt = 0;
    unsigned n_culled = 0, n_with_entries = 0, n_computed = 0, n_reduced = 0, n_reduced_empty = 0, n_reduced_rows = 0;
    __shared__ Outliers s_out;

    const int64_t NT = (int64_t)V.T * (V.T + 1) / 2;
    const int64_t nq = (NT - A.rank + A.world - 1) / A.world;
    // interleaved share of the tile sequence (see score_sweep_kernel)
    const int64_t m0 = 0, m1 = (nq - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x;
    auto tile_no = [&](int64_t m) { return ((int64_t)blockIdx.x + m * gridDim.x) * A.world + A.rank; };
    auto stage_cuts = [&](int ti, int tj, int b) {
        if (tid < TILE) {
            const int64_t gi = (int64_t)ti * TILE + tid, gj = (int64_t)tj * TILE + tid;
            c1I[b * TILE + tid] = A.cut1[gi];
            c1J[b * TILE + tid] = A.cut1[gj];
            c2I[b * TILE + tid] = l2on ? A.cut2[gi] : -INFINITY;
            c2J[b * TILE + tid] = l2on ? A.cut2[gj] : -INFINITY;
        }
    };
    TileStore ts;
    ts.init(sTS);
    // Scan-ahead: the tile sequence is filtered BEFORE anything is loaded.  A tile pair that holds no store entry and
    // fails the tile-level test against the largest cut of its two tiles (A.tcmax, one value per tile) is skipped
    // without touching its anchor-distance rows; the eight warps test eight consecutive candidates at a time.
    __shared__ unsigned char s_flag[2][SWW];
    int scan_round = 0;
    const bool scan_on = V.cull && V.is_metric && A.tcmax != nullptr;
    auto next_surviving = [&](int64_t start) -> int64_t {
        if (!scan_on) return start;
        for (int64_t base = start; base < m1; base += SWW, ++scan_round) {
            const int64_t cnd = base + warp;
            bool surv = false;
            if (cnd < m1) {
                int ci, cj;
                const int64_t t = tile_no(cnd);
                tile_from_index(t, V.T, ci, cj);
                surv = ci == cj || V.tl_ptr[t + 1] > V.tl_ptr[t] ||
                       tile_can_pass<1>(V, M, ci, cj, fmaxf(A.tcmax[ci], A.tcmax[cj]), nullptr);
            }
            if (lane == 0) s_flag[scan_round & 1][warp] = surv ? 1 : 0;
            __syncthreads();
            int first = -1;
#pragma unroll
            for (int w = SWW - 1; w >= 0; --w)
                if (s_flag[scan_round & 1][w]) first = w;
            if (first >= 0) {
                n_culled += first;
                ++scan_round;
                return base + first;
            }
            n_culled += (unsigned)min((int64_t)SWW, m1 - base);
        }
        return m1;
    };
    // m = the tile being computed, mnext = the one being prefetched, mnn = found while mnext's loads are in flight
    int64_t m = next_surviving(m0), mnext = m < m1 ? next_surviving(m + 1) : m1, mnn = m1;
    int k = 0;  // surviving tiles so far: buffer and descriptor slot
    if (m < m1) {
        const int64_t t = tile_no(m);
        const long long e0 = V.tl_ptr[t], e1 = V.tl_ptr[t + 1];
        if (tid == 0) {
            sDesc[0].base = e0;
            sDesc[0].end = e1;
        }
        int ti, tj;
        tile_from_index(t, V.T, ti, tj);
        load_point_tile(V, ti, sD0i, sM0i);
        load_point_tile(V, tj, sD0j, sM0j);
        load_tile_codes(V, e0, e1, sCode);
        stage_cuts(ti, tj, 0);
        cp_async_commit();
    }
    for (; m < m1; m = mnext, mnext = mnn, ++k) {
        const int buf = k & 1;
        int ti, tj;
        tile_from_index(tile_no(m), V.T, ti, tj);
        cp_async_wait_all();
        __syncthreads();
        if (mnext < m1) {
            int ni, nj;
            const int64_t t = tile_no(mnext);
            tile_from_index(t, V.T, ni, nj);
            const long long e0 = V.tl_ptr[t], e1 = V.tl_ptr[t + 1];
            if (tid == 0) {
                sDesc[(k + 1) & 3].base = e0;
                sDesc[(k + 1) & 3].end = e1;
            }
            load_point_tile(V, ni, buf ? sD0i : sD1i, buf ? sM0i : sM1i);
            load_point_tile(V, nj, buf ? sD0j : sD1j, buf ? sM0j : sM1j);
            load_tile_codes(V, e0, e1, sCode + (buf ^ 1) * TL_CAP);
            stage_cuts(ni, nj, buf ^ 1);
            cp_async_commit();
        }
        mnn = mnext < m1 ? next_surviving(mnext + 1) : m1;
        const float *sDi = buf ? sD1i : sD0i, *sDj = buf ? sD1j : sD0j;
        const PointMeta *sMi = buf ? sM1i : sM0i, *sMj = buf ? sM1j : sM0j;
        const float *k1I = c1I + buf * TILE, *k1J = c1J + buf * TILE;
        const float *k2I = c2I + buf * TILE, *k2J = c2J + buf * TILE;
        // tile-level pruning: no store entry in the tile and even the smallest possible prediction
        // exceeds every cut of the two tiles (phase 1 keeps a pair iff pred <= max(cut_i, cut_j))
        const bool has_entries = sDesc[k & 3].end != sDesc[k & 3].base;
        bool reduced = false;
        if (V.cull && V.is_metric && ti != tj) {
            const float cutmax = fmaxf(fmaxf(tile_max128(k1I), tile_max128(k2I)), fmaxf(tile_max128(k1J), tile_max128(k2J)));
            const bool can = tile_can_pass<1>(V, M, ti, tj, cutmax, nullptr);
            if (!can && !has_entries) {
                ++n_culled;
                continue;
            }
            if (A.reduced) {  // reduced tile mode (sweep.cuh): which rows can pass through their own cut?
                if (can) {
                    const int l = tid & (TILE - 1);
                    const float myc = tid < TILE ? fmaxf(k1I[l], k2I[l]) : fmaxf(k1J[l], k2J[l]);
                    // ... and, for those, the sharper test of the point itself against the other tile
                    const bool mine = tile_can_pass<1>(V, M, ti, tj, myc, nullptr) && tid < 2 * TILE &&
                                      (tid < TILE ? point_can_pass<1>(V, M, sDi, l, sMi[l].cA, tj, myc, nullptr)
                                                  : point_can_pass<1>(V, M, sDj, l, sMj[l].cA, ti, myc, nullptr));
                    reduced = collect_outliers(&s_out, mine);
                } else {
                    no_outliers(&s_out);
                    reduced = true;
                }
                // one pair per thread is ~3x less efficient than the vectorised tile: not worth it beyond ~3000 pairs
                if (reduced && (s_out.n_i + s_out.n_j) * TILE + (long long)(sDesc[k & 3].end - sDesc[k & 3].base) > REDUCED_MAX_ITEMS)
                    reduced = false;
            }
        }
        if (has_entries) ++n_with_entries;
        if (reduced) ++n_reduced;
        else ++n_computed;
        build_tile_store(V, ts, &sDesc[k & 3], sCode + buf * TL_CAP, sBm, 0);
        const uint32_t *bm = sBm;
        // ---- phase 2: drain the warp's queue, one survivor per lane ----
        auto drain = [&](int qn) {
            for (int e0 = 0; e0 < qn; e0 += 32) {
                const int e = e0 + lane;
                if (e < qn) {
                    const Survivor s = queue[e];
                    const int li2 = s.ids & 0xff, lj = s.ids >> 8;
                    const int gi = ti * TILE + li2, gj = tj * TILE + lj;
                    if (gj < V.n && s.lb < INFINITY) {
                        const PointMeta pi = sMi[li2], pj = sMj[lj];
                        if (is_candidate(pi, pj)) {
                            const PairVal pv = pair_value(V, ts, tm, M, s.lb, s.ub, li2, lj, gi, gj, li2, lj, pi, pj, sDi,
                                                          sDj, flag_bit(bm, li2, lj));
                            const float v = pv.v;
                            const uint32_t cflag = pv.computed ? 0x80000000u : 0u;
                            if (v <= k1I[li2] || (!pv.computed && v <= k2I[li2])) {
                                const int pos = atomicAdd(&A.cnt[gi], 1);
                                if (pos < A.R) A.rec[(int64_t)gi * A.R + pos] = make_uint2(__float_as_uint(v), (uint32_t)gj | cflag);
                            }
                            if (v <= k1J[lj] || (!pv.computed && v <= k2J[lj])) {
                                const int pos = atomicAdd(&A.cnt[gj], 1);
                                if (pos < A.R) A.rec[(int64_t)gj * A.R + pos] = make_uint2(__float_as_uint(v), (uint32_t)gi | cflag);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) *qcnt = 0;
            __syncwarp();
        };
        if (reduced) {
            // outlier rows / columns and store entries only, one pair per thread and round (sweep.cuh)
            const TileDesc *dp = &sDesc[k & 3];
            const int n_items = (s_out.n_i + s_out.n_j) * TILE + (int)(dp->end - dp->base);
            if (n_items == 0) ++n_reduced_empty;
            n_reduced_rows += s_out.n_i + s_out.n_j;
            for (int base = 0; base < n_items; base += SWT) {
                int li = 0, lj = 0;
                bool keep = false;
                Survivor sv;
                if (base + tid < n_items && reduced_item(V, &s_out, dp, sCode + buf * TL_CAP, base + tid, li, lj)) {
                    bounds_pair(sDi, sDj, na, li, lj, sv.lb, sv.ub);
                    const float s2 = sDi[sMj[lj].cA * SROW + li] + sDj[sMi[li].cA * SROW + lj];
                    int bin;
                    const float y = predict_clip2(tm, M, sv.lb, sv.ub, s2, bin);
                    keep = flag_bit(bm, li, lj) || y <= fmaxf(fmaxf(k1I[li], k2I[li]), fmaxf(k1J[lj], k2J[lj]));
                }
                const unsigned mm = __ballot_sync(0xffffffffu, keep);
                int qn = *reinterpret_cast<volatile int *>(qcnt);
                if (keep) {
                    sv.ids = (uint32_t)li | ((uint32_t)lj << 8);
                    sv.pad = 0;
                    queue[qn + __popc(mm & ((1u << lane) - 1u))] = sv;
                }
                qn += __popc(mm);
                __syncwarp();
                if (lane == 0) *qcnt = qn;
                __syncwarp();
                if (qn > A.qcap - 32) drain(qn);
            }
            drain(*reinterpret_cast<volatile int *>(qcnt));
            continue;
        }
        float cj[8];
        int cAj[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int lj = micro_off(tx, c);
            cj[c] = fmaxf(k1J[lj], k2J[lj]);
            cAj[c] = sMj[lj].cA * SROW;
        }
        float lb[4][8], ub[4][8];
#pragma unroll 1
        for (int pass = 0; pass < SW_PASSES; ++pass) {  // per half pass: 4 row steps + 1 drain-only step
            const int h = pass >= 5 ? 1 : 0, step = pass - 5 * h;
            const int row0 = h * (SWT / 4) + ty * 4;
            if (step == 0) {
                bounds_half(sDi, sDj, na, row0, tx, lb, ub);
                if (ti == tj) mask_diagonal<true>(row0, tx, lb, ub);
            }
            auto row_step = [&](const float (&lbr)[8], const float (&ubr)[8], int r) {
                const int li = row0 + r;
                const float ci = fmaxf(k1I[li], k2I[li]);
                const float *dj_row = sDj + sMi[li].cA * SROW;
                uint32_t w0, w1;
                flag_words(bm, li, tx, w0, w1);
                uint32_t km = (w0 & 0xfu) | ((w1 & 0xfu) << 4);  // flagged pairs always go to phase 2
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float s2 = sDi[cAj[c] + li] + dj_row[micro_off(tx, c)];
                    int bin;
                    const float y = predict_clip2(tm, M, lbr[c], ubr[c], s2, bin);
                    km |= (y <= fmaxf(ci, cj[c])) ? (1u << c) : 0u;
                }
                stage_row(queue, qcnt, lbr, ubr, km, li, tx);
            };
            switch (step) {
                case 0: row_step(lb[0], ub[0], 0); break;
                case 1: row_step(lb[1], ub[1], 1); break;
                case 2: row_step(lb[2], ub[2], 2); break;
                case 3: row_step(lb[3], ub[3], 3); break;
                default: break;
            }
            __syncwarp();
            const int qn = *reinterpret_cast<volatile int *>(qcnt);
            if (pass < SW_PASSES - 1 && qn <= A.qcap - QROW) continue;
            drain(qn);
        }
    }
    if (A.counters && tid == 0) {
        atomicAdd(&A.counters[0], (unsigned long long)n_culled);
        atomicAdd(&A.counters[1], (unsigned long long)n_with_entries);
        atomicAdd(&A.counters[2], (unsigned long long)n_computed);
        atomicAdd(&A.counters[3], (unsigned long long)n_reduced);
        atomicAdd(&A.counters[4], (unsigned long long)n_reduced_empty);
        atomicAdd(&A.counters[5], (unsigned long long)n_reduced_rows);
    }
}

// per tile: the largest of max(cutA, cutB) over its 128 points (what tile_max128 gives a sweep for the staged
// cuts); one warp per tile.  Lets a sweep test a tile pair before it loads anything of it.
__global__ void __launch_bounds__(256)
tile_cutmax_kernel(const float *__restrict__ cutA, const float *__restrict__ cutB, int T, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int t = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
    if (t >= T) return;
    float m = -INFINITY;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t g = (int64_t)t * TILE + q * 32 + lane;
        m = fmaxf(m, cutB ? fmaxf(cutA[g], cutB[g]) : cutA[g]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) out[t] = m;
}

int launch_tile_cutmax(annb_ctx *c, const float *cutA, const float *cutB, int T, float *out)
{
    ANNB_LAUNCH(tile_cutmax_kernel, (T * 32 + 255) / 256, 256, 0, c->stream, cutA, cutB, T, out);
    return ANNB_OK;
}

int launch_thresh_pairs(annb_ctx *c, ThreshPairArgs &A)
{
    const size_t lim = 227 * 1024;
    const size_t base = (size_t)4 * A.V.na * SROW * 4 + 4 * TILE * sizeof(PointMeta) + TS_BYTES +
                        2 * TL_CAP * 4 + 4 * sizeof(TileDesc) + 8 * TILE * 4 + sizeof(TileModel) + SWW * 4 + 64 +
                        sizeof(Outliers) + 64;  // (+ the static shared variables of the kernel)
    ANNB_REQUIRE(base + (size_t)SWW * (QROW + 32) * sizeof(Survivor) <= lim, ANNB_ERANGE,
                 "threshold pair sweep needs %zu bytes of shared memory (n_anchors=%d)",
                 base + (size_t)SWW * (QROW + 32) * sizeof(Survivor), A.V.na);
    int qcap = (int)((lim - base) / (SWW * sizeof(Survivor))) / 32 * 32;
    if (qcap > QCAP) qcap = QCAP;
    A.qcap = qcap;
    const size_t smem = base + (size_t)SWW * qcap * sizeof(Survivor);
    ANNB_CUDA(cudaFuncSetAttribute(thresh_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t NT = (int64_t)A.V.T * (A.V.T + 1) / 2;
    const int64_t nq = (NT - A.rank + A.world - 1) / A.world;
    int grid = c->num_sms;
    if (nq < grid) grid = nq < 1 ? 1 : (int)nq;
    ANNB_LAUNCH(thresh_pairs_kernel, grid, SWT, smem, c->stream, A);
    return ANNB_OK;
}

// Per row: the k1 smallest values and the k2 smallest not-computed (value, id) pairs of its records
// (n_src record lists of R slots per row: n_src = 1 after the pair sweep; world when merging the
// per-rank partial lists).  One warp per row, lists in shared memory, order-independent result.
__global__ void __launch_bounds__(256)
thresh_select_kernel(const uint2 *__restrict__ rec, const int32_t *__restrict__ cnt, int R, int64_t n, int k1,
                     int k2, int n_src, float *__restrict__ thresh, float *__restrict__ l1out,
                     float *__restrict__ l2val, int32_t *__restrict__ l2id)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *L1 = reinterpret_cast<float *>(smem) + warp * (k1 + 2 * k2 + 2);
    float *L2v = L1 + k1;
    int32_t *L2i = reinterpret_cast<int32_t *>(L2v + k2);
    float *thr = reinterpret_cast<float *>(L2i + k2);  // [0] k1-th value, [1] k2-th value
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < n; row += (int64_t)gridDim.x * 8) {
        for (int k = lane; k < k1; k += 32) L1[k] = INFINITY;
        for (int k = lane; k < k2; k += 32) {
            L2v[k] = INFINITY;
            L2i[k] = INT32_MAX;
        }
        if (lane == 0) {
            thr[0] = INFINITY;
            thr[1] = k2 > 0 ? INFINITY : -INFINITY;
        }
        __syncwarp();
        for (int src = 0; src < n_src; ++src) {
            const int64_t rrow = (int64_t)src * n + row;
            const int m = min(cnt[rrow], R);
            const uint2 *rr = rec + rrow * R;
            for (int e0 = 0; e0 < m; e0 += 32) {
                const int e = e0 + lane;
                float v = INFINITY;
                int32_t id = 0;
                bool comp = true, l2only = false;
                if (e < m) {
                    const uint2 x = rr[e];
                    v = __uint_as_float(x.x);
                    id = (int32_t)(x.y & 0x3fffffffu);
                    comp = (x.y >> 31) != 0;           // computed pair: not a guarantee_nmin candidate
                    l2only = (x.y & 0x40000000u) != 0; // merge input that is already in a value list
                }
                unsigned m1 = __ballot_sync(0xffffffffu, e < m && !l2only && v <= thr[0]);
                __syncwarp();  // every lane has read thr[0] before an insert rewrites it
                while (m1) {
                    const int s = __ffs(m1) - 1;
                    m1 &= m1 - 1;
                    const float xv = __shfl_sync(0xffffffffu, v, s);
                    if (xv <= thr[0]) SortedRow::insert(L1, nullptr, k1, xv, 0, lane, &thr[0]);
                }
                if (k2 > 0) {
                    unsigned m2 = __ballot_sync(0xffffffffu, e < m && !comp && v < INFINITY && v <= thr[1]);
                    __syncwarp();
                    while (m2) {
                        const int s = __ffs(m2) - 1;
                        m2 &= m2 - 1;
                        const float xv = __shfl_sync(0xffffffffu, v, s);
                        const int xi = __shfl_sync(0xffffffffu, id, s);
                        if (xv <= thr[1]) SortedRow::insert(L2v, L2i, k2, xv, xi, lane, &thr[1]);
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) thresh[row] = L1[k1 - 1];
        if (l1out)
            for (int k = lane; k < k1; k += 32) l1out[row * k1 + k] = L1[k];
        for (int k = lane; k < k2; k += 32) {
            l2val[row * k2 + k] = L2v[k];
            l2id[row * k2 + k] = L2i[k] == INT32_MAX ? -1 : L2i[k];
        }
        __syncwarp();
    }
}

int launch_thresh_select(annb_ctx *c, const uint2 *rec, const int32_t *cnt, int R, int64_t n, int k1, int k2,
                         int n_src, float *thresh, float *l1out, float *l2val, int32_t *l2id)
{
    const size_t smem = (size_t)8 * (k1 + 2 * k2 + 2) * 4;
    const int grid = (int)std::min<int64_t>((n + 7) / 8, (int64_t)c->num_sms * 8);
    ANNB_LAUNCH(thresh_select_kernel, grid, 256, smem, c->stream, rec, cnt, R, n, k1, k2, n_src, thresh, l1out, l2val,
                l2id);
    return ANNB_OK;
}

}  // namespace annb
