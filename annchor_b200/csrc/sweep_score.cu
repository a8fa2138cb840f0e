// sweep_score.cu -- K2+K3 fused scoring sweep over the upper-triangular tiles, plus the cheap
// sub-sampling sweep that feeds the sampler.
//
// Scoring (annchor/annchor.py:416-457, annchor/utils.py:581-589): for every not-computed
// candidate pair   p = max(thresh[i], thresh[j]) - RefineApprox,
//                  prob = searchsorted(errs[label], p, 'left') / len(errs[label]).
// prob takes at most sum(len+1) distinct values; the host ranks them once ("levels"), the
// sweep histograms the levels >= floor exactly and emits those pairs, and the top n_refine are
// then cut out of the emitted list -- the streaming equivalent of the reference's two
// np.argpartition calls over a materialised prob array.
//
// Phase 1 keeps a pair only if  lower bound < max(thresh[i], thresh[j]) - min_label efloor,
// which is necessary for its level to reach the floor (RefineApprox >= lower bound).
#include "sweep.cuh"
#include "sweep_args.cuh"

namespace annb {

constexpr int EMIT_CAP = SWT == 512 ? 128 : 256;  // per-warp emission staging (entries)


// flush a warp's staged emissions with ONE global atomic
__device__ __forceinline__ void flush_emit(const ScoreArgs &A, uint64_t *ek, uint16_t *el, int &en, int lane)
{
    if (en == 0) return;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(&A.counters[0], (unsigned long long)en);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (int k = lane; k < en; k += 32)
        if (base + k < A.emit_cap) {
            A.emit_key[base + k] = ek[k];
            A.emit_lvl[base + k] = el[k];
        }
    __syncwarp();
    en = 0;
}

// Tile processing: phase 1 runs over the 8 micro-tile rows and appends survivors to the warp's
// shared-memory queue; phase 2 (ONE code instance: the row steps and the end-of-tile flush are
// cases of a rolled 9-step loop) drains the queue in full 32-lane batches when the tile is done,
// or earlier if the next row step could overflow it.
__global__ void __launch_bounds__(SWT, 1) score_sweep_kernel(const __grid_constant__ ScoreArgs A)
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
    unsigned char *sTS = reinterpret_cast<unsigned char *>(sM1j + TILE);  // store state of the current tile
    uint32_t *sBm = reinterpret_cast<uint32_t *>(sTS);                    // its flag bitmap
    uint32_t *sCode = reinterpret_cast<uint32_t *>(sTS + TS_BYTES);       // [2][TL_CAP] staged store codes
    TileDesc *sDesc = reinterpret_cast<TileDesc *>(sCode + 2 * TL_CAP);   // [4] descriptors in flight
    unsigned char *sp = reinterpret_cast<unsigned char *>(sDesc + 4);
    Survivor *queue = reinterpret_cast<Survivor *>(sp) + warp * A.qcap;
    sp += (size_t)SWW * A.qcap * sizeof(Survivor);
    uint64_t *ek = reinterpret_cast<uint64_t *>(sp) + warp * EMIT_CAP;  // per-warp emission staging
    sp += SWW * EMIT_CAP * sizeof(uint64_t);
    uint16_t *el = reinterpret_cast<uint16_t *>(sp) + warp * EMIT_CAP;
    sp += SWW * EMIT_CAP * sizeof(uint16_t);
    float *thI = reinterpret_cast<float *>(sp);  // [2][128] thresh of the row tile (double-buffered)
    float *thJ = thI + 2 * TILE;                 // [2][128] thresh of the column tile
    TileModel *tm = reinterpret_cast<TileModel *>(thJ + 2 * TILE);
    int *qcnt = reinterpret_cast<int *>(tm + 1) + warp;  // [SWW] queue lengths
    uint32_t *sHist = reinterpret_cast<uint32_t *>(reinterpret_cast<int *>(tm + 1) + SWW);
    float *sErr = reinterpret_cast<float *>(sHist + A.nlevels);       // error tables + level ranks, if they fit
    uint16_t *sRank = reinterpret_cast<uint16_t *>(sErr + A.n_errs);
    const float *errs = A.tables_in_smem ? sErr : A.errs;
    const uint16_t *ranktab = A.tables_in_smem ? sRank : A.ranktab;
    build_tile_model(M, tm);
    if (tid < MAX_BINS) {
        // the error label of a pair is its regression bin b, or b + 1 when 2*dad sits exactly on edge
        // b + 1 (closed [lo, hi] labels vs (lo, hi] bins): mx = margins of (b, b+1) for level >= floor
        // (.x, .y) and for level > floor (.z, .w); en = that edge
        const int t1 = tid + 1 < M.nb ? tid + 1 : (M.nb > 0 ? M.nb - 1 : 0);
        const bool live = tid < M.nb, off = A.ef_min == -INFINITY;  // off: floor 0 or non-metric
        tm->mx[tid] = make_float4(off ? -INFINITY : (live ? A.efloor[tid] : INFINITY),
                                  off ? -INFINITY : (live ? A.efloor[t1] : INFINITY),
                                  off ? -INFINITY : (live ? A.efloor_hi[tid] : INFINITY),
                                  off ? -INFINITY : (live ? A.efloor_hi[t1] : INFINITY));
        tm->en[tid] = (tid + 1 < M.nb) ? M.e2[tid + 1] : INFINITY;
    }
    if (lane == 0) *qcnt = 0;
    for (int k = tid; k < A.nlevels; k += blockDim.x) sHist[k] = 0;
    if (A.tables_in_smem) {
        for (int k = tid; k < A.n_errs; k += blockDim.x) sErr[k] = A.errs[k];
        for (int k = tid; k < A.n_errs + M.nb; k += blockDim.x) sRank[k] = A.ranktab[k];
    }

    // this CTA's share of the rank's tile sequence, interleaved (CTA b takes b, b + grid, b + 2 grid, ...):
    // with spatially ordered points the expensive tiles sit around the diagonal, i.e. in runs of the
    // sequence, and a contiguous slice per CTA would leave most SMs idle while a few finish
    const int64_t nq = (A.q_end - A.q_begin + A.q_stride - 1) / A.q_stride;
    const int64_t m0 = 0, m1 = (nq - (int64_t)blockIdx.x + gridDim.x - 1) / gridDim.x;
    unsigned n_nc = 0, n_swept = 0, n_staged = 0, n_flagged = 0;  // per lane: < 2^32 (widened when reduced)
    int n_culled = 0, n_reduced = 0;
    __shared__ Outliers s_out;
    int en = 0;  // staged emissions of this warp

    auto tile_no = [&](int64_t m) {
        return (A.q_begin + ((int64_t)blockIdx.x + m * gridDim.x) * A.q_stride) * A.world + A.rank;
    };
    auto tile_of = [&](int64_t m, int &ti, int &tj) { tile_from_index(tile_no(m), V.T, ti, tj); };
    TileStore ts;
    ts.init(sTS);
    // Scan-ahead (see thresh_pairs_kernel): the tile sequence is filtered before anything is loaded -- no store
    // entry and the tile-level test fails for the largest threshold of the two tiles (A.tcmax) => skipped without
    // touching the tile's anchor-distance rows; eight candidates are tested at a time, one per warp.
    __shared__ unsigned char s_flag[2][SWW];
    int scan_round = 0;
    const bool scan_on = V.cull && A.ef_min > -INFINITY && A.tcmax != nullptr;
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
                       tile_can_pass<0>(V, M, ci, cj, fmaxf(A.tcmax[ci], A.tcmax[cj]), A.efloor);
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
            n_culled += (int)min((int64_t)SWW, m1 - base);
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
        if (tid < TILE) {
            thI[tid] = A.thresh[(int64_t)ti * TILE + tid];
            thJ[tid] = A.thresh[(int64_t)tj * TILE + tid];
        }
        cp_async_commit();
    }
    for (; m < m1; m = mnext, mnext = mnn, ++k) {
        const int buf = k & 1;
        int ti, tj;
        tile_of(m, ti, tj);
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
            if (tid < TILE) {
                thI[(buf ^ 1) * TILE + tid] = A.thresh[(int64_t)ni * TILE + tid];
                thJ[(buf ^ 1) * TILE + tid] = A.thresh[(int64_t)nj * TILE + tid];
            }
            cp_async_commit();
        }
        mnn = mnext < m1 ? next_surviving(mnext + 1) : m1;
        const float *sDi = buf ? sD1i : sD0i, *sDj = buf ? sD1j : sD0j;
        const PointMeta *sMi = buf ? sM1i : sM0i, *sMj = buf ? sM1j : sM0j;
        const float *tI = thI + buf * TILE, *tJ = thJ + buf * TILE;
        // tile-level pruning: no store entry in the tile and even the smallest possible prediction
        // cannot reach the floor level for the largest threshold of the two tiles (the same test
        // `cut - pred > margin` phase 1 applies per pair, at the tile's extremes)
        const bool has_entries = sDesc[k & 3].end != sDesc[k & 3].base;
        bool reduced = false;
        if (V.cull && A.ef_min > -INFINITY && ti != tj) {
            const float cutmax = fmaxf(tile_max128(tI), tile_max128(tJ));
            const bool can = tile_can_pass<0>(V, M, ti, tj, cutmax, A.efloor);
            if (!can && !has_entries) {
                n_culled += 1;
                continue;
            }
            if (A.reduced) {  // reduced tile mode (sweep.cuh): which rows can reach the floor level through their own threshold?
                if (can) {
                    const int l = tid & (TILE - 1);
                    const float myc = tid < TILE ? tI[l] : tJ[l];
                    // ... and, for those, the sharper test of the point itself against the other tile
                    const bool mine = tile_can_pass<0>(V, M, ti, tj, myc, A.efloor) && tid < 2 * TILE &&
                                      (tid < TILE ? point_can_pass<0>(V, M, sDi, l, sMi[l].cA, tj, myc, A.efloor)
                                                  : point_can_pass<0>(V, M, sDj, l, sMj[l].cA, ti, myc, A.efloor));
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
        build_tile_store(V, ts, &sDesc[k & 3], sCode + buf * TL_CAP, sBm, 0);
        const uint32_t *bm = sBm;
        // ---- phase 2: drain the warp's queue, one survivor per lane ----
        auto drain = [&](int qn) {
            if (lane == 0) n_staged += qn;
            for (int e0 = 0; e0 < qn; e0 += 32) {
                if (en > EMIT_CAP - 32) flush_emit(A, ek, el, en, lane);
                const int e = e0 + lane;
                int lvl = -1;
                uint64_t key = 0;
                if (e < qn) {
                    const Survivor s = queue[e];
                    const int li2 = s.ids & 0xff, lj = s.ids >> 8;
                    const int gi = ti * TILE + li2, gj = tj * TILE + lj;
                    if (gj < V.n && s.lb < INFINITY) {  // gi < gj by construction (diagonal masked to +inf)
                        const PointMeta pi = sMi[li2], pj = sMj[lj];
                        if (is_candidate(pi, pj)) {
                            const bool fl = flag_bit(bm, li2, lj);
                            n_flagged += fl ? 1 : 0;
                            const PairVal pv = pair_value(V, ts, tm, M, s.lb, s.ub, li2, lj, gi, gj, li2, lj, pi, pj,
                                                          sDi, sDj, fl);
                            if (!pv.computed) {
                                ++n_nc;
                                const float p = fmaxf(tI[li2], tJ[lj]) - pv.v;
                                const int label = err_label2(M, pv.dad);
                                if (p > A.efloor[label]) {
                                    int l = A.floor_level;
                                    if (p > A.efloor_hi[label]) {  // above the floor level: rank it
                                        const float *er = errs + M.eoff[label];
                                        int lo = 0, hi = M.eoff[label + 1] - M.eoff[label];
                                        while (lo < hi) {  // np.searchsorted(errs[label], p, 'left')
                                            const int mid = (lo + hi) >> 1;
                                            if (er[mid] < p) lo = mid + 1;
                                            else hi = mid;
                                        }
                                        l = ranktab[M.eoff[label] + label + lo];
                                    }
                                    if (l >= A.floor_level) {
                                        key = pair_key((uint32_t)gi, (uint32_t)gj);
                                        if (l > A.floor_level ||
                                            tie_key((uint32_t)gi, (uint32_t)gj, A.tie_salt) <= A.floor_mix_thr) {
                                            lvl = l;
                                            atomicAdd(&sHist[l], 1u);  // at the floor level: emitted pairs only
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
                if (A.emit_key) {
                    const unsigned mm = __ballot_sync(0xffffffffu, lvl >= 0);
                    if (lvl >= 0) {
                        const int pos = en + __popc(mm & ((1u << lane) - 1));
                        ek[pos] = key;
                        el[pos] = (uint16_t)lvl;
                    }
                    en += __popc(mm);
                    __syncwarp();
                }
            }
            __syncwarp();
            if (lane == 0) *qcnt = 0;
            __syncwarp();
        };
        if (reduced) {
            // outlier rows / columns and store entries only, one pair per thread and round; the per-pair test is the
            // one phase 1 of the full tile applies (same expressions, same rounding)
            const TileDesc *dp = &sDesc[k & 3];
            const int n_items = (s_out.n_i + s_out.n_j) * TILE + (int)(dp->end - dp->base);
            const uint32_t thr_hi_r = (uint32_t)(A.floor_mix_thr >> 32);
            for (int base = 0; base < n_items; base += SWT) {
                int li = 0, lj = 0;
                bool keep = false;
                Survivor sv;
                if (base + tid < n_items && reduced_item(V, &s_out, dp, sCode + buf * TL_CAP, base + tid, li, lj)) {
                    n_swept += 1;
                    bounds_pair(sDi, sDj, na, li, lj, sv.lb, sv.ub);
                    const float s2 = sDi[sMj[lj].cA * SROW + li] + sDj[sMi[li].cA * SROW + lj];
                    int bin;
                    const float y = predict_clip2(tm, M, sv.lb, sv.ub, s2, bin);
                    const float4 mx = tm->mx[bin];
                    const bool eq = s2 == tm->en[bin];
                    const float cut = fmaxf(tI[li], tJ[lj]);
                    if (!A.wide_floor) {
                        keep = cut - y > (eq ? mx.y : mx.x);
                    } else {
                        const uint32_t h = hash_pair32((uint32_t)(ti * TILE + li), (uint32_t)(tj * TILE + lj),
                                                       (uint32_t)A.tie_salt);
                        keep = (cut - y > (eq ? mx.w : mx.z)) || ((cut - y > (eq ? mx.y : mx.x)) && h <= thr_hi_r);
                    }
                    keep = keep || flag_bit(bm, li, lj);
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
            n_reduced += 1;
            continue;
        }
        // ---- phase 1: bounds + clipped prediction; keep (pred < cut-off) | flagged ----
        float cj[8];
        int cAj[8];
        uint32_t hj[8];  // column part of the pair hash (hash_pair32)
        const uint32_t thr_hi = (uint32_t)(A.floor_mix_thr >> 32);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            cj[c] = tJ[micro_off(tx, c)];
            cAj[c] = sMj[micro_off(tx, c)].cA * SROW;
            hj[c] = (uint32_t)(tj * TILE + micro_off(tx, c)) * 0x85EBCA77u;
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
                const float ci = tI[li];
                const float *dj_row = sDj + sMi[li].cA * SROW;
                uint32_t w0, w1;
                flag_words(bm, li, tx, w0, w1);
                uint32_t km = (w0 & 0xfu) | ((w1 & 0xfu) << 4);  // flagged pairs always go to phase 2
                // level >= floor needs  max(th_i, th_j) - RefineApprox > efloor[label]
                if (!A.wide_floor) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float s2 = sDi[cAj[c] + li] + dj_row[micro_off(tx, c)];
                        int bin;
                        const float y = predict_clip2(tm, M, lbr[c], ubr[c], s2, bin);
                        const float4 mx = tm->mx[bin];
                        const float mg = (s2 == tm->en[bin]) ? mx.y : mx.x;
                        km |= (fmaxf(ci, cj[c]) - y > mg) ? (1u << c) : 0u;  // the very expression phase 2 tests
                    }
                } else {
                    // many more pairs tie AT the floor level than the cut can take: of those, only the
                    // ones whose tie-break hash can pass the emission threshold go on (levels above the
                    // floor always do)
                    const uint32_t hi_row = ((uint32_t)(ti * TILE + li) * 0x9E3779B1u) ^ (uint32_t)A.tie_salt;
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float s2 = sDi[cAj[c] + li] + dj_row[micro_off(tx, c)];
                        int bin;
                        const float y = predict_clip2(tm, M, lbr[c], ubr[c], s2, bin);
                        const float4 mx = tm->mx[bin];
                        const bool eq = s2 == tm->en[bin];
                        const float cut = fmaxf(ci, cj[c]);
                        uint32_t h = hi_row ^ hj[c];
                        h ^= h >> 15;
                        h *= 0x2C1B3C6Du;
                        h ^= h >> 12;
                        h *= 0x297A2D39u;
                        h ^= h >> 15;
                        const bool above = cut - y > (eq ? mx.w : mx.z);  // same arithmetic as phase 2: p = cut - v
                        const bool atfl = (cut - y > (eq ? mx.y : mx.x)) && h <= thr_hi;
                        km |= (above | atfl) ? (1u << c) : 0u;
                    }
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
            drain(qn);
        }
        n_swept += 32 * SW_HALVES;
    }
    if (A.emit_key) flush_emit(A, ek, el, en, lane);
    __syncthreads();
    for (int k = tid; k < A.nlevels; k += blockDim.x)
        if (sHist[k]) atomicAdd(&A.hist[k], sHist[k]);
    unsigned long long w_nc = n_nc, w_swept = n_swept, w_flagged = n_flagged;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        w_nc += __shfl_xor_sync(0xffffffffu, w_nc, o);
        w_swept += __shfl_xor_sync(0xffffffffu, w_swept, o);
        w_flagged += __shfl_xor_sync(0xffffffffu, w_flagged, o);
    }
    if (lane == 0) {
        atomicAdd(&A.counters[1], w_nc);
        atomicAdd(&A.counters[2], w_swept);
        atomicAdd(&A.counters[3], (unsigned long long)n_staged);  // phase-1 survivors handed to phase 2
        atomicAdd(&A.counters[4], w_flagged);  // of which carried a flag bit (known / tightened / forced)
        if (warp == 0) {
            atomicAdd(&A.counters[5], (unsigned long long)n_culled);   // tiles skipped by the tile-level bound
            atomicAdd(&A.counters[6], (unsigned long long)n_reduced);  // tiles computed in reduced mode
        }
    }
}

int launch_score_sweep(annb_ctx *c, ScoreArgs &A)
{
    const size_t lim = 227 * 1024;
    const size_t base = (size_t)4 * A.V.na * SROW * 4 + 4 * TILE * sizeof(PointMeta) +
                        TS_BYTES + 2 * TL_CAP * 4 + 4 * sizeof(TileDesc) +
                        (size_t)SWW * EMIT_CAP * 10 + 4 * TILE * 4 + sizeof(TileModel) +
                        SWW * 4 + (size_t)A.nlevels * 4 + 64;
    const size_t tables = ((size_t)A.n_errs * 4 + (size_t)(A.n_errs + A.M.nb) * 2 + 15) & ~(size_t)15;
    const size_t qmin = (size_t)SWW * (QROW + 32) * sizeof(Survivor);
    ANNB_REQUIRE(base + qmin <= lim, ANNB_ERANGE,
                 "score sweep needs %zu bytes of shared memory (n_anchors=%d, %d levels)", base + qmin,
                 A.V.na, A.nlevels);
    A.tables_in_smem = base + qmin + tables <= lim ? 1 : 0;
    size_t avail = lim - base - (A.tables_in_smem ? tables : 0);
    int qcap = (int)(avail / (SWW * sizeof(Survivor))) / 32 * 32;
    if (qcap > QCAP) qcap = QCAP;
    A.qcap = qcap;
    const size_t smem = base + (A.tables_in_smem ? tables : 0) + (size_t)SWW * qcap * sizeof(Survivor);
    ANNB_CUDA(cudaFuncSetAttribute(score_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    const int64_t nq = (A.q_end - A.q_begin + A.q_stride - 1) / A.q_stride;
    int grid = c->num_sms;
    if (nq < grid) grid = nq < 1 ? 1 : (int)nq;
    ANNB_LAUNCH(score_sweep_kernel, grid, SWT, smem, c->stream, A);
    return ANNB_OK;
}

// ---------------------------------------------------------------------------------------------
// Sampler sweep (annchor/samplers.py:75-140, annchor/utils.py:543-578): emit a uniform
// sub-sample of the not-computed candidate pairs with their dad.  A pair participates iff
// hash32(pair, seed) <= thr -- thr = 0xffffffff enumerates every pair (exact mode, small
// problems).  No anchor loop: ~a dozen issue slots per pair; phase 2 handles the sub-sample.
// ---------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256, 2) sample_sweep_kernel(const SampleArgs A)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const View &V = A.V;
    const int na = V.na;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ty = tid >> 4, tx = tid & 15;
    float *sDi = reinterpret_cast<float *>(smem);
    float *sDj = sDi + na * SROW;
    PointMeta *sMi = reinterpret_cast<PointMeta *>(sDj + na * SROW);
    PointMeta *sMj = sMi + TILE;
    uint32_t *queue = reinterpret_cast<uint32_t *>(sMj + TILE) + warp * QCAP;  // li | lj << 8
    const int64_t NT = (int64_t)V.T * (V.T + 1) / 2;
    const int64_t nq = (NT - A.rank + A.world - 1) / A.world;
    const int64_t per = (nq + gridDim.x - 1) / gridDim.x;
    const int64_t m0 = (int64_t)blockIdx.x * per, m1 = min(nq, m0 + per);
    for (int64_t m = m0; m < m1; ++m) {
        // two-stage uniform sub-sample: tiles first (point ids carry no geometry, so a tile is a random
        // block of pairs), then pairs inside the visited tiles
        const uint64_t tix = (uint64_t)(m * A.world + A.rank);
        if (A.tile_thr != 0xffffffffu &&
            hash_pair32((uint32_t)tix, (uint32_t)(tix >> 32), A.seed ^ 0x5bd1e995u) > A.tile_thr)
            continue;
        int ti, tj;
        tile_from_index(m * A.world + A.rank, V.T, ti, tj);
        __syncthreads();
        load_point_tile(V, ti, sDi, sMi);
        load_point_tile(V, tj, sDj, sMj);
        cp_async_commit();
        cp_async_wait_all();
        __syncthreads();
        const bool diag = ti == tj;
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int li = micro_off(ty, r);
            const uint32_t gi = (uint32_t)(ti * TILE + li);
            uint32_t km = 0;  // columns of this row step that take part
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int lj = micro_off(tx, c);
                const uint32_t gj = (uint32_t)(tj * TILE + lj);
                const uint32_t h = hash_pair32(gi, gj, A.seed);
                bool keep = h <= A.thr && (!diag || li < lj) && gj < (uint32_t)V.n;
                if (keep && A.nb > 0) {
                    // stratified mode: A.thr is the largest per-bin threshold (cheap pre-filter); the pair's own
                    // threshold follows from its sampler bin [lo, hi) (utils.py:547-549) on the double anchor distance
                    const float dad = 0.5f * (sDi[sMj[lj].cA * SROW + li] + sDj[sMi[li].cA * SROW + lj]);
                    int b = 0;
                    for (int k = 1; k < A.nb; ++k) b += dad >= A.edge[k];
                    keep = A.bthr[b] != 0u && h <= A.bthr[b];
                }
                km |= keep ? (1u << c) : 0u;
            }
            // compact the kept (row, column) pairs of the warp into its queue: exclusive scan of the lanes' counts
            const int mine = __popc(km);
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const int qn = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t *qr = queue + (r & 1) * (QCAP / 2);  // row steps alternate between the halves (<= 256 entries each)
            int pos = incl - mine;
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if ((km >> c) & 1u) qr[pos++] = (uint32_t)li | ((uint32_t)micro_off(tx, c) << 8);
            warp_barrier();
            for (int e0 = 0; e0 < qn; e0 += 32) {
                const int e = e0 + lane;
                bool emit = false;
                uint64_t key = 0;
                float dad = 0.0f;
                if (e < qn) {
                    const uint32_t ids = qr[e];
                    const int li2 = ids & 0xff, lj = ids >> 8;
                    const int gi2 = ti * TILE + li2, gj = tj * TILE + lj;
                    const PointMeta pi = sMi[li2], pj = sMj[lj];
                    if (pi.slot < 0 && pj.slot < 0 && is_candidate(pi, pj)) {  // anchor pairs are computed
                        // only the hash-selected sub-sample gets here: the known-pair check goes to the hash
                        // map directly, so the per-tile lists need not be current for this sweep
                        float a, b;
                        const bool known = hash_lookup(V, pair_key((uint32_t)gi2, (uint32_t)gj), a, b) == KIND_KNOWN;
                        if (!known) {
                            emit = true;
                            key = pair_key((uint32_t)gi2, (uint32_t)gj);
                            dad = 0.5f * (sDi[pj.cA * SROW + li2] + sDj[pi.cA * SROW + lj]);
                        }
                    }
                }
                const unsigned mm = __ballot_sync(0xffffffffu, emit);
                if (mm) {
                    unsigned long long base = 0;
                    if (lane == 0) base = atomicAdd(A.counter, (unsigned long long)__popc(mm));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    const unsigned long long pos = base + __popc(mm & ((1u << lane) - 1));
                    if (emit && pos < A.out_cap) {
                        A.out_key[pos] = key;
                        A.out_dad[pos] = dad;
                    }
                }
            }
            warp_barrier();
        }
    }
}

int launch_sample_sweep(annb_ctx *c, const SampleArgs &A)
{
    const size_t smem = (size_t)2 * A.V.na * SROW * 4 + 2 * TILE * sizeof(PointMeta) +
                        (size_t)8 * QCAP * 4 + 64;
    ANNB_CUDA(cudaFuncSetAttribute(sample_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   (int)smem));
    const int64_t NT = (int64_t)A.V.T * (A.V.T + 1) / 2;
    const int64_t nq = (NT - A.rank + A.world - 1) / A.world;
    int grid = c->num_sms * 2;
    if (nq < grid) grid = nq < 1 ? 1 : (int)nq;
    ANNB_LAUNCH(sample_sweep_kernel, grid, 256, smem, c->stream, A);
    return ANNB_OK;
}

}  // namespace annb
