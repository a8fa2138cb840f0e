/*
 * annb.h -- C ABI of libannb.so, the B200 (sm_100a) implementation of the
 * `Annchor.fit()` hot path of gchq/annchor (reference v1.1.0).
 *
 * The reference has no FFI: its hot path is numba/joblib Python.  Each entry
 * point below replaces one Python-level function or plug on that path and cites
 * it (paths relative to the upstream repository root).  INTEGRATION.md shows the
 * ctypes stub a reference maintainer would add for each.
 *
 * Conventions
 *   - every function returns 0 on success, a negative ANNB_E* code otherwise;
 *     annb_last_error() returns a thread-local message for the last failure.
 *   - no exception crosses the boundary; no torch / C++ types in signatures.
 *   - the caller owns every host buffer and passes pre-allocated outputs; the
 *     library owns device objects behind opaque handles.
 *   - a context is bound to one CUDA device and one stream and is NOT thread
 *     safe (the reference drives fit() from a single Python thread as well).
 *   - host arrays use the reference's dtypes (int64 indices, float64 values,
 *     C order) so numpy arrays pass through unchanged.
 *   - "hd" pointers are host pointers unless the function name ends in _dev.
 */
#ifndef ANNB_H
#define ANNB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANNB_OK 0
#define ANNB_EINVAL (-1)   /* bad argument */
#define ANNB_ECUDA (-2)    /* CUDA runtime error (message has the cudaError string) */
#define ANNB_ENOMEM (-3)   /* device or host allocation failed */
#define ANNB_ENOGPU (-4)   /* no usable CUDA device */
#define ANNB_ESTATE (-5)   /* call sequence violated (e.g. select before set_model) */
#define ANNB_ERANGE (-6)   /* capacity exceeded (output list, string length, n_anchors > 64) */

/* metrics: annchor/utils.py:62-86 (get_function_from_input string table) */
#define ANNB_EUCLIDEAN 0     /* annchor/distances.py:8-13 */
#define ANNB_COSINE 1        /* scipy.spatial.distance.cosine, annchor/utils.py:14,67 */
#define ANNB_LEVENSHTEIN 2   /* annchor/distances.py:16-20 */
#define ANNB_WASSERSTEIN1D 3 /* annchor/utils.py:75-86 with cost |a-b| (1-D) */
#define ANNB_WASSERSTEIN 4   /* annchor/utils.py:75-86 with a general cost matrix (exact OT, <= 64 bins) */

#define ANNB_F32 0
#define ANNB_F64 1
#define ANNB_U8 2

typedef struct annb_ctx annb_ctx;
typedef struct annb_dataset annb_dataset;
typedef struct annb_index annb_index;

const char *annb_last_error(void);
int annb_version(void);
/* number of kernels this library has launched since load (for bench.py's gpu_launches) */
int64_t annb_launch_count(void);

/* ---- context ------------------------------------------------------------ */
int annb_ctx_create(int device, annb_ctx **out);
int annb_ctx_destroy(annb_ctx *ctx);
/* device buffers released by indexes / datasets are cached for reuse by later fits; this returns
 * the cached blocks to the driver (also done by annb_ctx_destroy) */
int annb_pool_trim(void);
int annb_sync(annb_ctx *ctx);
/* cudaStream_t of the context, as an integer (for torch.cuda.ExternalStream / events) */
int annb_ctx_stream(annb_ctx *ctx, uint64_t *stream);
/* device-timed region on the context stream: start, stop -> milliseconds */
int annb_timer_start(annb_ctx *ctx);
int annb_timer_stop(annb_ctx *ctx, float *ms);

/* ---- datasets: the `X` argument of Annchor(X, ...) (annchor/annchor.py:92-118) --- */
/* dense rows, dtype ANNB_F32 / ANNB_F64, row-major (n, d); on_device != 0 => X is a
 * device pointer that is copied (device to device). */
int annb_dataset_dense(annb_ctx *ctx, const void *X, int64_t n, int64_t d, int dtype,
                       int on_device, annb_dataset **out);
/* strings packed as bytes + offsets[n+1] (one byte per code point). */
int annb_dataset_strings(annb_ctx *ctx, const uint8_t *chars, const int64_t *offsets, int64_t n,
                         annb_dataset **out);
/* histograms (n, nbins), dtype ANNB_F32 / ANNB_F64 / ANNB_U8; stored as unit-mass CDFs. */
int annb_dataset_hist(annb_ctx *ctx, const void *H, int64_t n, int64_t nbins, int dtype,
                      annb_dataset **out);
/* a new data set with the items of `ds` in the given order (item q of the result = item order[q]),
 * copied device to device; no counterpart in the reference (its arrays live in host memory) */
int annb_dataset_gather(annb_ctx *ctx, const annb_dataset *ds, const int64_t *order, int64_t n,
                        annb_dataset **out);
/* histograms + a general (nbins, nbins) ground-cost matrix for ANNB_WASSERSTEIN (annchor/utils.py:75-86:
 * kantorovich(x, y, cost=M) -- zero bins dropped, unit mass, exact optimal transport); nbins <= 64 */
int annb_dataset_hist_cost(annb_ctx *ctx, const void *H, int64_t n, int64_t nbins, int dtype,
                           const double *cost, annb_dataset **out);
int annb_dataset_free(annb_dataset *ds);
int64_t annb_dataset_len(const annb_dataset *ds);

/* ---- K4: get_exact_ijs(f, X, IJ) (annchor/utils.py:110-177; plug at
 *      annchor/annchor.py:77-82,178-183) --------------------------------------- */
/* out[p] = metric(X[ij[p][0]], X[ij[p][1]]);  ij is (n,2) int64 C-order, out float64[n] */
int annb_pair_dists(annb_ctx *ctx, const annb_dataset *ds, int metric, const int64_t *ij,
                    int64_t n, double *out);
/* device-resident variant: i, j int32[n] and out float[n] are device pointers */
int annb_pair_dists_dev(annb_ctx *ctx, const annb_dataset *ds, int metric, const int32_t *i,
                        const int32_t *j, int64_t n, float *out);

/* query twin get_exact_query_ijs(f, X, Z, IJ) (annchor/utils.py:180-245, annchor/annchor.py:657-661):
 * out[p] = metric(X[ij[p][0]], Z[ij[p][1]]).  `both` is ONE data set holding the nx items of X followed
 * by the items of Z (so that both sides share the device layout of the metric kernels). */
int annb_pair_dists_query(annb_ctx *ctx, const annb_dataset *both, int metric, int64_t nx, const int64_t *ij,
                          int64_t n, double *out);

/* ---- K1: anchor pickers (annchor/pickers.py) ------------------------------- */
/* MaxMinAnchorPicker.get_anchors (annchor/pickers.py:18-52): `first` is the caller's
 * np.random.randint(nx) draw; A int64[na]; D float64 (n, na) row-major (the layout of the
 * reference's returned D.T).  Either output may be NULL. */
int annb_maxmin_anchors(annb_ctx *ctx, const annb_dataset *ds, int metric, int64_t na,
                        int64_t first, int64_t *A, double *D);
/* Selected / Random pickers (annchor/pickers.py:86-128): distances from given anchors. */
int annb_anchor_dists(annb_ctx *ctx, const annb_dataset *ds, int metric, const int64_t *A,
                      int64_t na, double *D);

/* ---- stage operators on explicit pair lists (reference-shaped, float64) ------- */
/* get_bounds_njit_ijs (annchor/utils.py:274-301): D (nx,na) row-major; out (n,2) */
int annb_bounds_ijs(annb_ctx *ctx, const int64_t *ij, int64_t n, const double *D, int64_t nx,
                    int64_t na, double *bounds);
/* get_dad_ijs (annchor/utils.py:355-380) */
int annb_dad_ijs(annb_ctx *ctx, const int64_t *ij, int64_t n, const double *D, int64_t nx,
                 int64_t na, double *dad);
/* update_bounds / get_bounds_alt (annchor/utils.py:304-352): per-point known lists as CSR
 * (kptr[nx+1], kids ascending per row, kds); out (n,2) = (lb, ub) */
int annb_update_bounds(annb_ctx *ctx, const int64_t *ij, int64_t n, const int64_t *kptr,
                       const int64_t *kids, const double *kds, int64_t nx, double *bounds);
/* SimpleStratifiedLinearRegression.predict (annchor/regressors.py:71-103) followed by the
 * clip of annchor/annchor.py:359-363: features (n,4) = [lb, ub, dad, is_anchor];
 * bins[nb+1]; coef (nb,3); icpt[nb]; pred_raw and pred_clipped float64[n] (either may be NULL) */
int annb_predict_stratified(annb_ctx *ctx, const double *features, int64_t n, const double *bins,
                            const double *coef, const double *icpt, int64_t nb, double *pred_raw,
                            double *pred_clipped);
/* SimpleStratifiedErrorRegression.predict (annchor/error_predictors.py:56-67) */
int annb_error_labels(annb_ctx *ctx, const double *feature, int64_t n, const double *bins,
                      int64_t nb, int64_t *labels);
/* get_probs (annchor/utils.py:581-589): errs = concatenated sorted tables, eptr[nl+1] */
int annb_probs(annb_ctx *ctx, const double *p, const int64_t *labels, int64_t n,
               const double *errs, const int64_t *eptr, int64_t nl, double *prob);
/* thresh loop (annchor/annchor.py:399-404): out[i] = k-th smallest (0-based) of
 * RA[row_pairs[row_ptr[i]:row_ptr[i+1]]] */
int annb_row_kth(annb_ctx *ctx, const double *RA, int64_t npairs, const int64_t *row_ptr,
                 const int64_t *row_pairs, int64_t nx, int64_t k, double *out);
/* get_nn (annchor/utils.py:383-429): ngi int64 (nx, nn-1), ngd float64 (nx, nn-1) */
int annb_get_nn(annb_ctx *ctx, int64_t nx, int64_t nn, const double *RA, const int64_t *ij,
                int64_t npairs, const int64_t *row_ptr, const int64_t *row_pairs,
                const uint8_t *not_computed, int64_t *ngi, double *ngd);

/* ---- streaming index: Annchor.fit() without Theta(N^2) state -------------------
 * The candidate pair set (annchor/annchor.py:208-256), the per-pair features
 * (:258-303), predictions (:345-380), labels (:382-393) and probabilities (:395-436)
 * are never stored; they are re-derived tile by tile from the anchor-distance matrix.
 * Exactly-evaluated pairs live in a device hash map + per-tile flag bitmap. */
typedef struct annb_index_params {
    int32_t n_anchors;    /* annchor/annchor.py:97 */
    int32_t n_neighbors;  /* :98 */
    int32_t locality;     /* :106 */
    int32_t loc_thresh;   /* :107 */
    int32_t loc_min;      /* :108,167-168 (already resolved / clipped by the caller) */
    int32_t is_metric;    /* :110 */
    int32_t rank;         /* tile shard: this index sweeps tiles t with t % world == rank */
    int32_t world;
} annb_index_params;

int annb_index_create(annb_ctx *ctx, const annb_dataset *ds, int metric,
                      const annb_index_params *params, annb_index **out);
int annb_index_destroy(annb_index *ix);
/* optional: size the known-pair store for n_pairs entries up front (exact evaluations, i.e.
 * p_work * N(N-1)/2 of annchor/annchor.py:100,440, plus look-ahead pairs that may be tightened) */
int annb_index_reserve_pairs(annb_index *ix, int64_t n_pairs);
/* stage 1: run the MaxMin picker into the index (or load caller-provided anchors) */
int annb_index_maxmin(annb_index *ix, int64_t first, int64_t *A);
int annb_index_set_anchors(annb_index *ix, const int64_t *A, int64_t nA, const double *D);
/* Spatial renumbering (no counterpart in the reference; results are those of the reference algorithm
 * on the relabelled data set).  order[new] = old: points sorted by (closest anchor, distance to it, id),
 * so that the 128-point tiles of the sweeps are geometrically coherent and whole tile pairs can be
 * pruned from per-tile bounds.  Use: annb_index_maxmin / set_anchors on an index over the original
 * data set, annb_index_spatial_order, annb_dataset_gather, annb_index_create on the gathered data set,
 * annb_index_adopt_anchors(new, old, order). */
int annb_index_spatial_order(annb_index *ix, int64_t *order);
int annb_index_adopt_anchors(annb_index *ix, annb_index *src, const int64_t *order);
/* copy D out as float64 (n, na) row-major */
int annb_index_get_D(annb_index *ix, double *D);
/* candidate set (get_locality): returns number of candidate pairs P and the number of rows
 * whose locality threshold had to be relaxed (annchor/utils.py:472-480) */
int annb_index_locality(annb_index *ix, int64_t *n_candidates, int64_t *n_relaxed);
/* sampler support (annchor/samplers.py:75-140, annchor/utils.py:543-578): a uniform pool of the
 * not-computed candidate pairs with their double-anchor distance.  If at most max_pool such pairs
 * exist the pool is ALL of them (*exact = 1: the host can sort it into the reference's IJs order
 * and reproduce the reference's sampler bit for bit); otherwise a hash-selected uniform sub-sample. */
int annb_index_sample_pool(annb_index *ix, uint64_t seed, int64_t max_pool, int64_t *n_pool,
                           int64_t *n_not_computed, int *exact);
/* stratified refill: replaces the pool by the not-computed candidates whose double anchor distance
 * falls into bin b = [bins[b], bins[b+1]) (annchor/utils.py:547-549), each kept with probability
 * rate[b] (0 skips the bin) -- used when the uniform pool holds too few pairs of a bin */
int annb_index_sample_pool_bins(annb_index *ix, uint64_t seed, const double *bins, const double *rate,
                                int64_t nb, int64_t max_pool, int64_t *n_pool);
/* copy the pool out: ij (n_pool, 2) int64, dad float64[n_pool] */
int annb_index_get_pool(annb_index *ix, int64_t *ij, double *dad);
/* features [lb, ub, dad] of explicit pairs in the sweeps' float32 arithmetic: feat (n, 3) */
int annb_index_pair_features(annb_index *ix, const int64_t *ij, int64_t n, double *feat);
/* the store's entry for explicit pairs: kind[p] = 0 none, 1 exactly known (a = distance),
 * 2 tightened by update_anchor_points ((a, b) = (lb, ub), annchor/annchor.py:503-510),
 * 3 forced by guarantee_nmin (annchor/utils.py:619).  a / b may be NULL. */
int annb_index_pair_state(annb_index *ix, const int64_t *ij, int64_t n, int32_t *kind, double *a,
                          double *b);
/* mark pairs as exactly computed with distance d (sample_y / refine results,
 * annchor/annchor.py:342,380,472-473) */
int annb_index_add_known(annb_index *ix, const int64_t *ij, const double *d, int64_t n);
/* evaluate the metric on pairs and mark them (get_exact_ijs + the above); d out may be NULL */
int annb_index_eval_pairs(annb_index *ix, const int64_t *ij, int64_t n, double *d);
/* regression / error model (annchor/regressors.py:39-67, error_predictors.py:26-54) */
int annb_index_set_model(annb_index *ix, const double *bins, const double *coef,
                         const double *icpt, int64_t nb, const double *errs, const int64_t *eptr);
/* thresh (annchor/annchor.py:399-404) over RefineApprox = exact where known else clipped
 * prediction; thresh float64[n] out (may be NULL; kept on device for select) */
int annb_index_row_thresh(annb_index *ix, double *thresh);
/* the thresholds of the last annb_index_row_thresh / annb_index_guarantee_nmin, float64[n] */
int annb_index_get_thresh(annb_index *ix, double *thresh);
/* guarantee_nmin (annchor/utils.py:606-621) with nmin = 3*nn//2; returns # forced pairs */
int annb_index_guarantee_nmin(annb_index *ix, int64_t nmin, int64_t *n_forced);
/* select_refine_candidate_pairs scoring + choice (annchor/annchor.py:416-465): picks the
 * n_refine most probable not-computed pairs (ties at the cut resolved by pair order) and the
 * following n_refine*(lookahead-1) as the look-ahead set.  Results stay on device;
 * *n_selected / *n_next report the counts. */
int annb_index_select(annb_index *ix, int64_t n_refine, int64_t lookahead, int64_t *n_selected,
                      int64_t *n_next);
/* copy the selected / look-ahead pairs to the host ((n,2) int64) */
int annb_index_get_selected(annb_index *ix, int64_t *ij_sel, int64_t *ij_next);
/* replace the look-ahead set (`nextback`, annchor/annchor.py:457,462) by caller-chosen pairs:
 * lets annb_index_update_bounds be driven -- and checked against the reference's
 * update_anchor_points -- on an explicit pair list */
int annb_index_set_lookahead(annb_index *ix, const int64_t *ij, int64_t n);
/* evaluate the metric on the selected pairs and mark them known (annchor/annchor.py:467-473) */
int annb_index_refine_selected(annb_index *ix, int64_t *n_evals);
/* update_anchor_points (annchor/annchor.py:475-512) on the look-ahead pairs */
int annb_index_update_bounds(annb_index *ix, int64_t *n_updated);
/* get_ann / get_nn (annchor/annchor.py:514-530): idx int64 (n, nn), dist float64 (n, nn),
 * column 0 = self / 0 */
int annb_index_neighbor_graph(annb_index *ix, int64_t *idx, double *dist);
/* Annchor.query (annchor/annchor.py:643-683, annchor/query_functions.py:183-212) against a fitted
 * index: `both` holds X followed by the nq query items; ngi int64 (nq, nn) / ngd float64 (nq, nn) are the
 * nn approximate nearest points of X per query; *n_evals = metric evaluations spent
 * (n_anchors * nq + n_refine).  p_work is the fraction of the nq * nx brute-force evaluations. */
int annb_index_query(annb_index *ix, const annb_dataset *both, int64_t nq, int64_t nn, double p_work,
                     int64_t *ngi, double *ngd, int64_t *n_evals);
/* counters: [0] pairs swept, [1] known pairs, [2] tightened pairs, [3] sweep launches,
 * [4] candidate pairs, [5] hash capacity, [6] anchor pairs, [7] not-computed candidates */
int annb_index_stats(annb_index *ix, int64_t *out, int64_t n);
/* device time (ms, CUDA events on the context stream) and pairs covered, summed over the full
 * scoring sweeps (score_sweep_kernel, stride 1) this index has run */
int annb_index_last_sweep(annb_index *ix, float *ms, int64_t *pairs);

/* ---- multi-GPU (no counterpart in the single-host reference) ----------------------
 * An index created with world > 1 sweeps only its shard of the tiles.  Small host-side reductions
 * (thresholds, level / digit histograms, branch flags) go through a caller-supplied sum all-reduce;
 * bulk per-rank results (evaluated pairs, tightened bounds) are exported device-to-device for the
 * caller's NCCL all-gather and imported on the other ranks. */
#define ANNB_RED_U64 0
#define ANNB_RED_F32 1
#define ANNB_RED_I32 2
#define ANNB_RED_DEVICE 0x100 /* or-ed into dtype: buf is a device pointer on the index's GPU */
/* in-place sum all-reduce over `count` elements of `dtype` (host memory unless ANNB_RED_DEVICE is
 * set; the library has synchronised its stream before the call); returns 0 on success */
typedef int (*annb_reduce_fn)(void *user, void *buf, int64_t count, int dtype);
int annb_index_set_reducer(annb_index *ix, annb_reduce_fn fn, void *user);
/* pairs this rank evaluated in the last annb_index_refine_selected: device int32 i, j and float d */
int annb_index_export_refined(annb_index *ix, int32_t *i_dev, int32_t *j_dev, float *d_dev,
                              int64_t cap, int64_t *n);
/* bounds this rank tightened in the last annb_index_update_bounds */
int annb_index_export_tightened(annb_index *ix, int32_t *i_dev, int32_t *j_dev, float *lb_dev,
                                float *ub_dev, int64_t cap, int64_t *n);
/* insert entries received from other ranks: kind 1 = known (a = d), 2 = tightened (a, b) = (lb, ub) */
int annb_index_import_dev(annb_index *ix, int kind, const int32_t *i_dev, const int32_t *j_dev,
                          const float *a_dev, const float *b_dev, int64_t n);

/* brute-force k-NN on device (recall oracle at sizes where the CPU cannot;
 * annchor/annchor.py:943-1023): idx int64 (n,k), dist float64 (n,k), column 0 = self */
int annb_bruteforce_knn(annb_ctx *ctx, const annb_dataset *ds, int metric, int64_t k,
                        int64_t *idx, double *dist);

/* exact nearest-enemy graph (annchor/annchor.py:685-786 approximates it from the fitted state): the nn nearest
 * items with a different label, idx int64 (n,nn), dist float64 (n,nn); -1 / +inf where fewer enemies exist */
int annb_nearest_enemies(annb_ctx *ctx, const annb_dataset *ds, int metric, const int32_t *labels, int64_t nn,
                         int64_t *idx, double *dist);

/* host helper: numba's in-@njit np.random.seed + np.random.choice(replace=False)
 * (annchor/utils.py:555-557,572) so the materialised sampler reproduces the reference's draw */
int annb_numba_rng_new(uint32_t seed, void **state);
int annb_numba_rng_free(void *state);
int annb_numba_rng_shuffle(void *state, int64_t *x, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* ANNB_H */
