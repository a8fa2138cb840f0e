// sweep_args.cuh -- kernel argument blocks of the tile sweeps (shared by sweep_*.cu and index.cu).
#pragma once
#include "index.cuh"

namespace annb {

struct ThreshArgs {
    View V;
    Model M;
    int k1;            // nn + 1
    int k2;            // nmin + 1 (0 = no second list)
    float *thresh;     // [npad]
    float *l2val;      // [n * k2]
    int32_t *l2id;     // [n * k2]
    int rank, world;
    int qcap;          // set by the launcher: per-warp survivor queue capacity
    int col_stride;    // visit the column tiles tc = col_phase, col_phase + col_stride, ... (1: all)
    int col_phase;
    int band;          // > 0: first the column tiles rb - band .. rb + band (spatially ordered index: that is
                       // where a row's near neighbours are), then the strided ones outside the band
    float *cut2;       // [npad] out (k2 > 0): k2-th smallest not-computed value of the visited columns
    const int32_t *rb_list;  // explicit row blocks to process (nullptr: all, split over the ranks)
    int n_rb;
};

// upper-triangle pass of the two-stage threshold computation: every pair that can be among the
// smallest values of either endpoint (value <= that endpoint's cut from the column-subset pre-pass)
// is appended to the endpoint's record list
struct ThreshPairArgs {
    View V;
    Model M;
    const float *cut1;   // [npad] k1-th smallest RefineApprox over a column subset (upper bound of thresh)
    const float *cut2;   // [npad] same for the not-computed list (nullptr: no second list)
    uint2 *rec;          // [n][R] records (value bits, other endpoint | computed << 31)
    int32_t *cnt;        // [n] records appended (may exceed R: the row is then recomputed the long way)
    int R;
    int rank, world;
    int qcap;
    unsigned long long *counters;  // trace (may be nullptr): [0] tiles pruned, [1] tiles with store entries, [2] tiles
                                   // computed in full, [3] tiles in reduced mode
    int reduced;         // reduced tile mode on (sweep.cuh)
    const float *tcmax;  // [T] largest cut of each tile (launch_tile_cutmax) or nullptr: no scan-ahead
};

struct ScoreArgs {
    View V;
    Model M;
    const float *thresh;      // [npad]
    const float *errs;        // concatenated sorted error tables (global, L1-resident)
    const uint16_t *ranktab;  // level of (label, r): ranktab[eoff[label] + label + r], r in [0, len]
    int n_errs;               // total entries of errs (ranktab has n_errs + nb)
    int tables_in_smem;       // set by the launcher: errs + ranktab are staged in shared memory
    int qcap;                 // set by the launcher: per-warp survivor queue capacity
    int nlevels;
    int floor_level;          // histogram / emit only levels >= floor_level
    uint64_t floor_mix_thr;   // AT the floor level emit only pairs with splitmix64(pair key ^ tie_salt) <= this
    uint64_t tie_salt;
    float efloor[MAX_BINS];   // p > efloor[label]  <=>  level(label, p) >= floor_level
    float efloor_hi[MAX_BINS];  // p > efloor_hi[label]  <=>  level(label, p) >= floor_level + 1
    int wide_floor;           // phase 1 pre-filters floor-level pairs on the tie-break hash (floor_mix_thr < all)
    float ef_min;             // min over labels of efloor (phase-1 margin); -inf disables the filter
    int has_forced;           // FORCED marks exist (iteration 0): flagged pairs always go to phase 2
    uint32_t *hist;           // [nlevels]
    unsigned long long *counters;  // [0] emitted, [1] not-computed candidates seen in phase 2, [2] pairs swept
    uint64_t *emit_key;
    uint16_t *emit_lvl;
    unsigned long long emit_cap;
    int64_t q_begin, q_end;   // range of this rank's tile sequence
    int q_stride;             // 1 = every tile; s > 1 = pilot over every s-th tile
    int rank, world;
    int reduced;              // reduced tile mode on (sweep.cuh)
    const float *tcmax;       // [T] largest threshold of each tile or nullptr: no scan-ahead
};

struct SampleArgs {
    View V;
    uint32_t thr;       // a pair of a visited tile takes part iff hash_pair32(i, j, seed) <= thr
    uint32_t tile_thr;  // a tile is visited iff hash of its index <= tile_thr (0xffffffff: every tile)
    uint32_t seed;
    // stratified mode (nb > 0): the pair's threshold depends on its dad bin [edge[b], edge[b+1]) --
    // bthr[b] = 0 skips the bin; thr / tile_thr are ignored, every tile is visited
    int nb;
    float edge[MAX_BINS + 1];
    uint32_t bthr[MAX_BINS];
    uint64_t *out_key;
    float *out_dad;
    unsigned long long out_cap;
    unsigned long long *counter;
    int rank, world;
};

int launch_thresh_sweep(annb_ctx *c, ThreshArgs &A);
int launch_thresh_pairs(annb_ctx *c, ThreshPairArgs &A);
int launch_tile_cutmax(annb_ctx *c, const float *cutA, const float *cutB, int T, float *out);
int launch_thresh_select(annb_ctx *c, const uint2 *rec, const int32_t *cnt, int R, int64_t n, int k1, int k2,
                         int n_src, float *thresh, float *l1out, float *l2val, int32_t *l2id);
int launch_score_sweep(annb_ctx *c, ScoreArgs &A);
int launch_sample_sweep(annb_ctx *c, const SampleArgs &A);

}  // namespace annb
