// index_obj.cuh -- the host-side index object (annb_index), shared by index.cu and query.cu.
#pragma once
#include <vector>

#include "index.cuh"

using namespace annb;

// ---------------------------------------------------------------------------------------------
// the index object
// ---------------------------------------------------------------------------------------------
struct annb_index {
    annb_ctx *ctx = nullptr;
    const annb_dataset *ds = nullptr;
    int metric = 0;
    annb_index_params P;
    int64_t n = 0, npad = 0;
    int na = 0, T = 0;
    int64_t NT = 0;  // upper-triangular tiles
    bool have_anchors = false, have_locality = false, have_model = false, have_thresh = false;
    bool has_forced = false;
    std::vector<int32_t> A_host;
    DevBuf A_dev, D64, D32, Dpm, meta, scratch;
    int dpitch = 32;
    // known-pair store
    DevBuf htab;
    // per-tile entry lists derived from the hash map (rebuilt lazily before a sweep)
    DevBuf tl_ptr, tl_cnt, tl_code, tl_a, tl_b, scan_tmp;
    // per-tile anchor-distance intervals and closest-anchor sets (tile-level pruning in the sweeps)
    DevBuf tb_cutmax;  // [T] largest cut per tile of the sweep about to run (scan-ahead)
    DevBuf tb_lo, tb_hi, tb_cm, order_dev;  // order_dev[new] = old id (int32) when the index is renumbered
    bool ordered = false;  // the points were renumbered by annb_index_spatial_order (coherent tiles)
    bool tl_dirty = false;
    int64_t tl_entries = 0;
    uint64_t hcap = 0;
    int64_t hcount_ub = 0;  // upper bound on occupied slots
    // model
    Model model;
    std::vector<float> errs_host;
    std::vector<uint16_t> rank_host;
    DevBuf errs_dev, rank_dev;
    int nlevels = 0;
    // candidate-set statistics (host copies, from annb_index_locality)
    std::vector<int32_t> ncand_host, anc_cand_host;
    int64_t n_candidates = 0, n_anchor_pairs = 0;
    // thresholds and row lists
    DevBuf thresh, l2val, l2id;
    // selection
    DevBuf hist, counters, emit_key, emit_lvl, sel_i, sel_j, nxt_i, nxt_j, tiehist, tiekeys;
    int64_t n_sel = 0, n_next = 0;
    // ties at a selection cut are broken by splitmix64(pair key ^ tie_salt); the salt changes with
    // every selection so that the tie-break of one iteration is independent of the previous ones
    // (the survivors of an earlier cut are exactly the pairs with LARGE mixed keys under its salt)
    uint64_t tie_salt = 0;
    int64_t n_selects = 0;
    // sampler pool
    DevBuf pool_key, pool_dad;
    int64_t n_pool = 0;
    // temporaries
    DevBuf t0, t1, t2, t3, t4, t5, t6;
    // CSR of known pairs
    DevBuf kptr, kids, kds, kdeg, gptr, gJ, gsrc, row_order, twork, theavy;
    // two-stage thresholds
    DevBuf tcut1, tcut2, trec, tcnt, tgat, tgcnt, l1part;
    int64_t csr_entries = 0;
    // stats
    int64_t pairs_swept = 0, sweeps = 0, n_tight = 0, n_known = 0;
    float last_sweep_ms = 0;
    int64_t last_sweep_pairs = 0;
    // multi-GPU: host-buffer sum all-reduce supplied by the caller (torch.distributed in Python)
    annb_reduce_fn reducer = nullptr;
    void *reducer_user = nullptr;
    // results of the last refine / tighten kept on device for export to the other ranks
    int64_t n_refined = 0, n_tightened = 0;

    int reduce(void *buf, int64_t count, int dtype) const
    {
        if (P.world <= 1) return ANNB_OK;
        if (!reducer) {
            set_error("index is sharded (world=%d) but no reducer was set (annb_index_set_reducer)", P.world);
            return ANNB_ESTATE;
        }
        const int rc = reducer(reducer_user, buf, count, dtype);
        if (rc != 0) {
            set_error("reducer callback failed with code %d", rc);
            return ANNB_ESTATE;
        }
        return ANNB_OK;
    }

    View view() const
    {
        View V;
        V.n = n;
        V.npad = npad;
        V.na = na;
        V.T = T;
        V.nn = P.n_neighbors;
        V.is_metric = P.is_metric;
        V.D32 = D32.as<float>();
        V.Dpm = Dpm.as<float>();
        V.dpitch = dpitch;
        V.meta = meta.as<PointMeta>();
        V.htab = htab.as<HashSlot>();
        V.hmask = hcap - 1;
        V.tl_ptr = tl_ptr.as<long long>();
        V.tl_code = tl_code.as<uint32_t>();
        V.tl_a = tl_a.as<float>();
        V.tl_b = tl_b.as<float>();
        V.tb_lo = tb_lo.as<float>();
        V.tb_hi = tb_hi.as<float>();
        V.tb_cm = tb_cm.as<uint64_t>();
        V.cull = cull_enabled ? 1 : 0;
        return V;
    }
    bool cull_enabled = true;
    bool reduced_enabled = true;  // reduced tile mode of the sweeps (sweep.cuh); ANNB_NO_REDUCED turns it off
    bool scan_enabled = true;     // scan-ahead over the tile sequence; ANNB_NO_SCAN turns it off
    int64_t n_not_computed() const { return n_candidates - n_anchor_pairs - n_known; }
};

