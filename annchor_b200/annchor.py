"""``Annchor(X, distance, ...).fit()`` / ``.neighbor_graph`` on the B200 -- the host-side mirror
of annchor/annchor.py:21-623 over the streaming device index (include/annb.h, annb_index_*).

Same constructor keywords, same stage order, same attribute names after ``fit()``
(``neighbor_graph``, ``A``, ``D``, ``evals``, ``n_samples``, ``p_work`` ...).  What differs, by
design, is that nothing Theta(N^2) is materialised: ``IJs``, ``features``, ``RefineApprox`` and
``not_computed_mask`` do not exist; the plug-ins see the protocols in ``annchor_b200.plugins``.
"""
import ctypes as C
import os
import time
from collections import Counter

import numpy as np

from . import _lib
from ._lib import check, ptr, as_c
from .core import Dataset, GpuExactIJs, default_context
from . import plugins
from .plugins import (FEATURE_NAMES, NothingToSample, MaxMinAnchorPicker, SimpleStratifiedSampler,
                      SimpleStratifiedLinearRegression, SimpleStratifiedErrorRegression,
                      regression_device_spec, error_device_spec)


REORDER_MIN_POINTS = 16384  # spatial renumbering from this size on (the two-stage thresholds start at 128 tiles)


class Index:
    """Thin wrapper over annb_index_* (one per fit)."""

    def __init__(self, ctx, dataset, n_anchors, n_neighbors, locality, loc_thresh, loc_min,
                 is_metric=True, rank=0, world=1):
        self._L = _lib.load()
        self.ctx = ctx
        self.dataset = dataset
        self.n = dataset.n
        self.na = n_anchors
        self.nn = n_neighbors
        p = _lib.IndexParams(n_anchors, n_neighbors, locality, loc_thresh, int(loc_min),
                             1 if is_metric else 0, rank, world)
        h = C.c_void_p()
        check(self._L.annb_index_create(ctx.handle, dataset.handle, dataset.metric, C.byref(p), C.byref(h)))
        self.handle = h
        self._n_nc = None

    def close(self):
        if getattr(self, "handle", None):
            self._L.annb_index_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reserve_pairs(self, n_pairs):
        check(self._L.annb_index_reserve_pairs(self.handle, int(n_pairs)))

    def maxmin(self, first):
        A = np.empty(self.na, dtype=np.int64)
        check(self._L.annb_index_maxmin(self.handle, int(first), ptr(A)))
        return A

    def set_anchors(self, A, D):
        A = as_c(A, np.int64)
        D = as_c(D, np.float64)
        if D.shape != (self.n, self.na):
            raise ValueError("D must have shape (nx, n_anchors) = %r, got %r" % ((self.n, self.na), D.shape))
        check(self._L.annb_index_set_anchors(self.handle, ptr(A), A.shape[0], ptr(D)))

    def spatial_order(self):
        """order[new] = old: points by (closest anchor, distance to it, id) (annb_index_spatial_order)."""
        order = np.empty(self.n, dtype=np.int64)
        check(self._L.annb_index_spatial_order(self.handle, ptr(order)))
        return order

    def adopt_anchors(self, src, order):
        order = as_c(order, np.int64)
        check(self._L.annb_index_adopt_anchors(self.handle, src.handle, ptr(order)))

    def get_D(self):
        D = np.empty((self.n, self.na), dtype=np.float64)
        check(self._L.annb_index_get_D(self.handle, ptr(D)))
        return D

    def locality(self):
        nc, nr = C.c_int64(), C.c_int64()
        check(self._L.annb_index_locality(self.handle, C.byref(nc), C.byref(nr)))
        return nc.value, nr.value

    def sample_pool(self, seed, max_pool):
        """-> (n_pool, n_not_computed, exact)"""
        n, nc, ex = C.c_int64(), C.c_int64(), C.c_int()
        check(self._L.annb_index_sample_pool(self.handle, int(seed) & 0xFFFFFFFFFFFFFFFF, int(max_pool),
                                             C.byref(n), C.byref(nc), C.byref(ex)))
        self._n_pool = n.value
        return n.value, nc.value, bool(ex.value)

    def sample_pool_bins(self, seed, bins, rate, max_pool):
        """Stratified refill (annb_index_sample_pool_bins) -> n_pool; read it with get_pool()."""
        bins, rate = as_c(bins, np.float64), as_c(rate, np.float64)
        n = C.c_int64()
        check(self._L.annb_index_sample_pool_bins(self.handle, int(seed) & 0xFFFFFFFFFFFFFFFF, ptr(bins), ptr(rate),
                                                  rate.shape[0], int(max_pool), C.byref(n)))
        self._n_pool = n.value
        return n.value

    def get_pool(self):
        ij = np.empty((self._n_pool, 2), dtype=np.int64)
        dad = np.empty(self._n_pool, dtype=np.float64)
        check(self._L.annb_index_get_pool(self.handle, ptr(ij), ptr(dad)))
        return ij, dad

    def pair_features(self, ij):
        ij = as_c(ij, np.int64).reshape(-1, 2)
        feat = np.empty((ij.shape[0], 3), dtype=np.float64)
        check(self._L.annb_index_pair_features(self.handle, ptr(ij), ij.shape[0], ptr(feat)))
        return feat

    def pair_state(self, ij):
        """-> (kind int32[n], a float64[n], b float64[n]): the store entry of each pair
        (0 none, 1 known: a = distance, 2 tightened: (a, b) = (lb, ub), 3 forced)."""
        ij = as_c(ij, np.int64).reshape(-1, 2)
        n = ij.shape[0]
        kind, a, b = np.zeros(n, dtype=np.int32), np.zeros(n), np.zeros(n)
        check(self._L.annb_index_pair_state(self.handle, ptr(ij), n, ptr(kind), ptr(a), ptr(b)))
        return kind, a, b

    def get_thresh(self):
        th = np.empty(self.n, dtype=np.float64)
        check(self._L.annb_index_get_thresh(self.handle, ptr(th)))
        return th

    def set_lookahead(self, ij):
        ij = as_c(ij, np.int64).reshape(-1, 2)
        check(self._L.annb_index_set_lookahead(self.handle, ptr(ij), ij.shape[0]))
        self._sel = (0, ij.shape[0])

    def add_known(self, ij, d):
        ij, d = as_c(ij, np.int64).reshape(-1, 2), as_c(d, np.float64)
        check(self._L.annb_index_add_known(self.handle, ptr(ij), ptr(d), ij.shape[0]))

    def eval_pairs(self, ij):
        ij = as_c(ij, np.int64).reshape(-1, 2)
        d = np.empty(ij.shape[0], dtype=np.float64)
        check(self._L.annb_index_eval_pairs(self.handle, ptr(ij), ij.shape[0], ptr(d)))
        return d

    def set_model(self, bins, coef, icpt, errs=None, eptr=None):
        bins, coef, icpt = as_c(bins, np.float64), as_c(coef, np.float64), as_c(icpt, np.float64)
        if errs is not None:
            errs, eptr = as_c(errs, np.float64), as_c(eptr, np.int64)
        check(self._L.annb_index_set_model(self.handle, ptr(bins), ptr(coef), ptr(icpt), bins.shape[0] - 1,
                                           ptr(errs), ptr(eptr)))

    def row_thresh(self):
        th = np.empty(self.n, dtype=np.float64)
        check(self._L.annb_index_row_thresh(self.handle, ptr(th)))
        return th

    def guarantee_nmin(self, nmin):
        nf = C.c_int64()
        check(self._L.annb_index_guarantee_nmin(self.handle, int(nmin), C.byref(nf)))
        return nf.value

    def select(self, n_refine, lookahead):
        a, b = C.c_int64(), C.c_int64()
        check(self._L.annb_index_select(self.handle, int(n_refine), int(lookahead), C.byref(a), C.byref(b)))
        self._sel = (a.value, b.value)
        return a.value, b.value

    def get_selected(self):
        ns, nx = self._sel
        s, x = np.empty((ns, 2), dtype=np.int64), np.empty((nx, 2), dtype=np.int64)
        check(self._L.annb_index_get_selected(self.handle, ptr(s), ptr(x)))
        return s, x

    def refine_selected(self):
        n = C.c_int64()
        check(self._L.annb_index_refine_selected(self.handle, C.byref(n)))
        return n.value

    def update_bounds(self):
        n = C.c_int64()
        check(self._L.annb_index_update_bounds(self.handle, C.byref(n)))
        return n.value

    def neighbor_graph(self):
        idx = np.empty((self.n, self.nn), dtype=np.int64)
        dist = np.empty((self.n, self.nn), dtype=np.float64)
        check(self._L.annb_index_neighbor_graph(self.handle, ptr(idx), ptr(dist)))
        return idx, dist

    def stats(self):
        out = np.zeros(8, dtype=np.int64)
        check(self._L.annb_index_stats(self.handle, ptr(out), 8))
        return dict(zip(["pairs_swept", "n_known", "n_tight", "sweeps", "n_candidates", "hash_capacity",
                         "n_anchor_pairs", "n_not_computed"], out.tolist()))

    def query(self, both, nq, nn, p_work):
        """annb_index_query: `both` = Dataset of X followed by the nq queries -> (ngi, ngd, n_evals)."""
        ngi = np.empty((nq, nn), dtype=np.int64)
        ngd = np.empty((nq, nn), dtype=np.float64)
        ne = C.c_int64()
        check(self._L.annb_index_query(self.handle, both.handle, int(nq), int(nn), float(p_work), ptr(ngi), ptr(ngd),
                                       C.byref(ne)))
        return ngi, ngd, ne.value

    def last_sweep(self):
        ms, pairs = C.c_float(), C.c_int64()
        check(self._L.annb_index_last_sweep(self.handle, C.byref(ms), C.byref(pairs)))
        return ms.value, pairs.value

    # -- multi-GPU hooks (annchor_b200.dist) ---------------------------------------------------
    def set_reducer(self, cb):
        self._reducer = cb  # keep the ctypes callback alive as long as the index
        check(self._L.annb_index_set_reducer(self.handle, C.cast(cb, C.c_void_p), None))

    def export_refined(self, i_dev, j_dev, d_dev, cap):
        n = C.c_int64()
        check(self._L.annb_index_export_refined(self.handle, i_dev, j_dev, d_dev, int(cap), C.byref(n)))
        return n.value

    def export_tightened(self, i_dev, j_dev, lb_dev, ub_dev, cap):
        n = C.c_int64()
        check(self._L.annb_index_export_tightened(self.handle, i_dev, j_dev, lb_dev, ub_dev, int(cap),
                                                  C.byref(n)))
        return n.value

    def import_dev(self, kind, i_dev, j_dev, a_dev, b_dev, n):
        check(self._L.annb_index_import_dev(self.handle, int(kind), i_dev, j_dev, a_dev, b_dev, int(n)))


class Annchor:
    """Quickly computes the approximate k-NN graph for slow metrics -- on a B200.

    Parameters are those of the reference class (annchor/annchor.py:27-88).  ``func`` must be one
    of the bundled metric names ('euclidean', 'cosine', 'levenshtein', 'wasserstein'): arbitrary
    Python callables cannot run on the device and there is no CPU fallback, so they raise.
    ``backend`` is accepted for signature compatibility.  ``get_exact_ijs`` must be None or an
    ``annchor_b200.GpuExactIJs`` (the metric is always evaluated by the CUDA kernels; a host
    evaluator would be a CPU fallback and raises instead of being silently ignored).

    Plug-ins: anchor pickers use the reference protocol unchanged; samplers may implement either
    ``sample_index(ann)`` (streaming, any N) or the reference's ``sample(features, ...)``
    (annchor/samplers.py:75-110), which is served from a materialised view of the not-computed
    candidate pairs and therefore limited to ``plugins.MATERIALISED_LIMIT`` such pairs;
    regression / error predictor must expose their fitted tables (see annchor_b200.plugins).

    Device limits (checked here, before any work is done): n_anchors <= 64,
    n_neighbors <= 42 (guarantee_nmin keeps 3*n_neighbors//2 + 1 <= 64 entries per row),
    loc_thresh <= 7, n_partitions <= 8.
    """

    def __init__(self, X, func, func_kwargs=None, n_anchors=20, n_neighbors=15, n_samples=5000,
                 p_work=0.1, anchor_picker=None, sampler=None, regression=None, error_predictor=None,
                 random_seed=42, locality=5, loc_thresh=1, loc_min=None, verbose=False, is_metric=True,
                 get_exact_ijs=None, backend="loky", niters=2, lookahead=5, device=0, ctx=None,
                 comm=None, _dataset=None, _trace=None):
        if not isinstance(func, str):
            raise NotImplementedError(
                "annchor_b200 evaluates metrics on the GPU; pass one of 'euclidean', 'cosine', "
                "'levenshtein', 'wasserstein' (got a callable).  There is no CPU fallback.")
        if func not in _lib.METRIC_IDS:
            raise AssertionError("Error: The string must be one of %s" % sorted(_lib.METRIC_IDS))
        cost = None
        if func == "wasserstein":
            assert func_kwargs is not None and "cost_matrix" in func_kwargs, \
                "Error: wassetstein metric requires cost_function kwarg"
            cost = func_kwargs["cost_matrix"]
        assert backend in ["loky", "multiprocessing"]
        if get_exact_ijs is not None and not isinstance(get_exact_ijs, GpuExactIJs):
            raise NotImplementedError(
                "get_exact_ijs=%r: annchor_b200 evaluates the metric with its CUDA kernels; a host-side "
                "evaluator (annchor/annchor.py:77-82) would be a CPU fallback, which this package does not "
                "provide.  Pass None or an annchor_b200.GpuExactIJs." % (get_exact_ijs,))
        if n_anchors > 64:
            raise ValueError("n_anchors=%d: the device keeps the nearest-anchor set in a 64-bit mask "
                             "(n_anchors <= 64)" % n_anchors)
        if 3 * n_neighbors // 2 + 1 > 64 or n_neighbors < 2:
            raise ValueError("n_neighbors=%d: the device row lists hold 64 entries and guarantee_nmin "
                             "(annchor/utils.py:606-621) needs 3*n_neighbors//2 + 1 of them "
                             "(2 <= n_neighbors <= 42)" % n_neighbors)
        if loc_thresh > 7:
            raise ValueError("loc_thresh=%d: the device counts shared anchors up to 7" % loc_thresh)
        self.X = X
        self.nx = len(X)
        self.N = (self.nx * (self.nx - 1)) // 2
        self.f = func
        self.evals = 0
        self.n_anchors = n_anchors
        self.na = int(np.sum([self.nx - j for j in range(1, self.n_anchors + 1)]))
        self.n_neighbors = n_neighbors
        self.p_work = p_work
        self.n_samples = n_samples
        # annchor/annchor.py:132-148
        if self.p_work > 1:
            print("Warning: p_work should not exceed 1.  Setting it to 1.")
            self.p_work = 1.0
        min_p_work = (2 * (self.na + self.n_samples) + 1) / self.N
        min_p_work = 1 if min_p_work > 1 else min_p_work
        if self.p_work < min_p_work:
            print("Warning: Too many anchors/samples for specified p_work.")
            print("Increasing p_work to %5.3f." % min_p_work)
            self.p_work = min_p_work
        if self.p_work > 0.75:
            print("Warning: High Value of p_work.")
            print("Think about decreasing n_anchors or n_samples, or using BruteForce.")
        self.anchor_picker = anchor_picker or MaxMinAnchorPicker()
        self.sampler = sampler or SimpleStratifiedSampler()
        self.regression = regression or SimpleStratifiedLinearRegression()
        self.error_predictor = error_predictor or SimpleStratifiedErrorRegression()
        if not hasattr(self.sampler, "sample_index"):
            if not hasattr(self.sampler, "sample"):
                raise NotImplementedError("samplers must implement sample_index(ann) or the reference "
                                          "protocol sample(features, feature_names, n_samples, "
                                          "not_computed_mask, random_seed)")
            # reference-protocol sampler (annchor/samplers.py:75-110): materialised view
            self.sampler = plugins.MaterialisedSamplerAdapter(self.sampler)
        self.random_seed = random_seed
        self.verbose = verbose
        self.locality = locality
        self.loc_thresh = loc_thresh
        self.loc_min = 10 * self.n_neighbors if loc_min is None else loc_min
        self.loc_min = int(np.clip(self.loc_min, 0, self.nx - 1))
        self.is_metric = is_metric
        self.niters = niters
        self.lookahead = lookahead
        self.backend = backend
        self.feature_names = list(FEATURE_NAMES)
        self.ctx = ctx or default_context(device)
        self._cost = cost
        # _dataset: an already-uploaded annchor_b200.Dataset of X (skips the host-to-device copy)
        self._dataset = _dataset if _dataset is not None else Dataset(self.ctx, X, func, cost_matrix=cost)
        self._plug = GpuExactIJs(func, self.ctx, cost)
        self._plug._cache = (X, self._dataset)
        self.get_exact_ijs = self._plug  # same call contract as annchor/annchor.py:77-82
        # comm: an annchor_b200.dist.Comm -- this process is one rank of a sharded fit (every rank
        # calls fit() with the same X and arguments; the tile sweeps are split, results identical)
        self.comm = comm
        rank, world = (comm.rank, comm.world) if comm is not None else (0, 1)
        self._index = Index(self.ctx, self._dataset, n_anchors, n_neighbors, locality, loc_thresh,
                            self.loc_min, is_metric, rank=rank, world=world)
        # size the known-pair store once: all exact evaluations + the look-ahead pairs that
        # update_anchor_points may tighten (annchor.py:440,444-457)
        budget = int(self.p_work * self.N)
        n_ref = max(budget // max(niters, 1), 0)
        if not (self.nx >= REORDER_MIN_POINTS and is_metric and os.environ.get("ANNB_NO_REORDER") is None):
            # (with spatial renumbering the store is sized on the index that replaces this one)
            self._index.reserve_pairs(min(budget + n_ref * (lookahead - 1) * max(niters - 1, 0) + n_samples * niters,
                                          self.N) + 1024)
        self._xchg = None
        if comm is not None and world > 1:
            from .dist import IndexExchange
            self._index.set_reducer(comm.reducer)
            self._xchg = IndexExchange(self._index, comm)
        self._D = None
        # Spatial renumbering (large problems): after the anchors are known the points are sorted by
        # (closest anchor, distance to it), the data set is gathered in that order on the device and the
        # fit runs on the relabelled points; results are mapped back.  Tiles of the sweeps are then
        # geometrically coherent and most tile pairs are pruned from per-tile bounds.
        self._reorder = (self.nx >= REORDER_MIN_POINTS and is_metric and os.environ.get("ANNB_NO_REORDER") is None)
        self._order = None
        self._index_args = (n_anchors, n_neighbors, locality, loc_thresh, self.loc_min, is_metric, rank, world)
        self._reserve = min(budget + n_ref * (lookahead - 1) * max(niters - 1, 0) + n_samples * niters, self.N) + 1024
        self.stage_times = {}
        # _trace: a dict that receives per-stage results (sample, thresholds, selected / look-ahead
        # sets, ...) -- used by the stage-parity tests; costs device-to-host copies, off by default
        self._trace = _trace
        self._it = 0

    # -- stage 1 -----------------------------------------------------------------------------
    def get_anchors(self):
        A, D, evals = self.anchor_picker.get_anchors(self)
        self.A = np.asarray(A)
        if D is not None:  # host picker (reference protocol): hand its distances to the index
            self._index.set_anchors(self.A.astype(np.int64) if self.A.size else np.zeros(0, np.int64),
                                    np.asarray(D, dtype=np.float64))
        self.evals += evals
        if self._reorder:
            self._renumber()

    def _renumber(self):
        """Replace the index by one over the points in spatial order (device-side gather)."""
        order = self._index.spatial_order()
        ds2 = self._dataset.gather(order)
        ix2 = Index(self.ctx, ds2, *self._index_args)
        ix2.reserve_pairs(self._reserve)
        ix2.adopt_anchors(self._index, order)
        if self._xchg is not None:
            from .dist import IndexExchange
            ix2.set_reducer(self.comm.reducer)
            self._xchg = IndexExchange(ix2, self.comm)
        self._index.close()
        self._index, self._dataset_ordered, self._order = ix2, ds2, order

    def _to_user_ids(self, ij):
        return ij if self._order is None else self._order[ij]

    @property
    def D(self):
        if self._D is None:
            D = self._index.get_D()
            if self._order is not None:
                Du = np.empty_like(D)
                Du[self._order] = D
                D = Du
            self._D = D
        return self._D

    def get_locality(self):
        self.n_candidates, self.n_relaxed = self._index.locality()

    # -- sample / model ----------------------------------------------------------------------
    def get_sample(self):
        ijs, feats, bins = self.sampler.sample_index(self)
        self.sample_ijs = ijs
        self.sample_bins = bins
        self.n_samples = ijs.shape[0]
        self.sample_features = np.hstack([feats, np.zeros((feats.shape[0], 1))])
        self.sample_y = self._index.eval_pairs(ijs)  # evaluates the metric and marks the pairs computed
        self.evals += self.sample_y.shape[0]
        if self._trace is not None:
            self._trace["sample_ijs%d" % self._it] = self._to_user_ids(ijs)
            self._trace["sample_bins%d" % self._it] = np.array(bins, copy=True)
            self._trace["sample_features%d" % self._it] = self.sample_features.copy()

    def fit_predict_regression(self):
        self.regression.fit(self.sample_features, self.feature_names, self.sample_y,
                            sample_bins=self.sample_bins)
        self.sample_predict = self.regression.predict(self.sample_features, self.feature_names)
        if self._trace is not None:
            self._trace["coef%d" % self._it] = regression_device_spec(self.regression, self.feature_names)[1]

    def fit_predict_errors(self):
        self.error_predictor.fit(self.sample_features, self.feature_names,
                                 self.sample_y - self.sample_predict, sample_bins=self.sample_bins)
        bins, coef, icpt = regression_device_spec(self.regression, self.feature_names)
        errs, eptr = error_device_spec(self.error_predictor, n_bins=bins.shape[0] - 1)
        self._index.set_model(bins, coef, icpt, errs, eptr)

    # -- select / refine ---------------------------------------------------------------------
    def select_refine_candidate_pairs(self, w=0.5, it=0):
        nn = self.n_neighbors
        if it == 0:
            self.n_forced = self._index.guarantee_nmin(3 * nn // 2)  # also computes thresh
        else:
            self._index.row_thresh()
        n_refine = int((self.p_work * self.N - self.na - self.n_samples) * w) + 1
        n_refine = 0 if n_refine < 0 else n_refine
        self.n_refine = n_refine
        self._index.select(n_refine, self.lookahead)
        if self._trace is not None:
            sel, nxt = self._index.get_selected()
            th = self._index.get_thresh()
            if self._order is not None:
                tu = np.empty_like(th)
                tu[self._order] = th
                th = tu
            self._trace["thresh%d" % it] = th
            self._trace["selected%d" % it] = self._to_user_ids(sel)
            self._trace["next%d" % it] = self._to_user_ids(nxt)
            if it == 0:
                self._trace["n_forced"] = self.n_forced
        n_eval = self._index.refine_selected()
        if self._xchg is not None:  # every rank evaluated its own share: make all stores identical
            self._xchg.refined()
            n_eval = self.comm.all_reduce_sum(n_eval)
        self.evals += n_eval

    def update_anchor_points(self):
        self.n_tightened = self._index.update_bounds()
        if self._xchg is not None:
            self._xchg.tightened()
            self.n_tightened = self.comm.all_reduce_sum(self.n_tightened)
        if self._trace is not None:
            self._trace["n_tightened%d" % self._it] = self.n_tightened

    def get_ann(self):
        # (a renumbered index reports the graph in the caller's numbering: annb_index_neighbor_graph)
        self.neighbor_graph = self._index.neighbor_graph()

    def to_sparse_matrix(self):
        """The k-NN graph as a dictionary-of-keys sparse distance matrix (annchor/annchor.py:625-641):
        D[i, j] = D[j, i] = distance + the smallest positive float64, for every (i, j) of the graph."""
        from scipy.sparse import coo_matrix
        idx, dist = self.neighbor_graph
        n, k = idx.shape
        eps = np.nextafter(0, 1, dtype=np.float64)
        rows = np.repeat(np.arange(n, dtype=np.int64), k)
        cols = idx.ravel()
        vals = dist.ravel() + eps
        ok = cols >= 0
        r = np.concatenate([rows[ok], cols[ok]])
        c = np.concatenate([cols[ok], rows[ok]])
        v = np.concatenate([vals[ok], vals[ok]])
        # one entry per (row, col): the reference's loop overwrites, a symmetric metric writes equal values
        _, first = np.unique(r * np.int64(n) + c, return_index=True)
        return coo_matrix((v[first], (r[first], c[first])), shape=(n, n)).todok()

    def get_nearest_enemies(self, y, nn=3, loc_min=100):
        """Nearest-enemy graph (annchor/annchor.py:685-786): for every point the ``nn`` nearest points carrying a
        different label, stored in ``self.nearest_enemy_graph = (ngi int64 (nx, nn), ngd float64 (nx, nn))``.

        The reference approximates this graph from its materialised candidate state (enemy locality ``loc_min``,
        regression predictions, the 50 best candidates per row evaluated exactly).  The streaming index keeps no
        per-pair state, so the graph is computed EXACTLY here, by exhaustive evaluation with the metric kernels
        (annb_nearest_enemies) -- an O(nx^2) pass meant for the data-set sizes the reference's own O(nx^2)-memory
        implementation can handle.  ``loc_min`` is accepted for signature compatibility."""
        y = np.asarray(y)
        nx = self.nx
        assert len(y) == nx, "Label dimension mismatch: len(y)=%d, len(X)=%d" % (len(y), nx)
        labels, inv, counts = np.unique(y, return_inverse=True, return_counts=True)
        assert len(labels) > 1, "Data must have more than one label"
        assert np.all(counts >= nn), "At least one label occurs fewer times than specified nn=%d" % nn
        self.nearest_enemy_graph = self._dataset.nearest_enemies(inv.astype(np.int32), nn)
        self.evals += nx * (nx - 1) // 2
        return self.nearest_enemy_graph

    def alpha_rss(self, y, dne=None, alpha=0):
        """annchor/annchor.py:921-940: greedy relaxed selective subset -- visit the points by increasing nearest
        enemy distance and keep a point unless an already kept point is closer than dne / (1 + alpha).  The
        distances to the kept points are evaluated on the device in blocks (same decisions as the reference's
        one-point-at-a-time loop: a block member only ever needs the kept points and earlier block members)."""
        if dne is None:
            if not hasattr(self, "nearest_enemy_graph"):
                self.get_nearest_enemies(y)
            dne = self.nearest_enemy_graph[1][:, 0]
        dne = np.asarray(dne, dtype=np.float64)
        ix = np.argsort(dne)
        alpha_dne = dne / (1 + alpha)
        rss = [int(ix[0])]
        self.rssDs = {}
        B = 256
        for b0 in range(0, len(ix), B):
            blk = ix[b0:b0 + B]
            old = np.array(rss, dtype=np.int64)
            both = np.concatenate([old, blk])
            IJ = np.stack([np.repeat(blk, both.shape[0]), np.tile(both, blk.shape[0])], 1)
            Dm = self._dataset.pair_dists(IJ).reshape(blk.shape[0], both.shape[0])
            self.evals += IJ.shape[0]
            pos_in_both = {int(v): old.shape[0] + q for q, v in enumerate(blk)}
            cols = list(range(old.shape[0]))  # columns of Dm that are kept points
            for q, i in enumerate(blk):
                ds = Dm[q, cols]
                self.rssDs[int(i)] = ds
                dnn = np.min(ds)
                if (dnn > alpha_dne[i]) or np.isclose(dnn, alpha_dne[i]):
                    rss.append(int(i))
                    cols.append(pos_in_both[int(i)])
        return np.array(rss)

    def annchor_selective_subset(self, y, dne=None, alpha=0):
        """annchor/annchor.py:788-919.  Not available: the reference's algorithm walks, for every point, the sorted
        list of ALL its candidate pairs with their RefineApprox / upper-bound values -- materialised per-pair state
        the streaming index does not keep.  ``alpha_rss`` (the greedy relaxed selective subset built on exact metric
        evaluations) and ``get_nearest_enemies`` are available."""
        raise NotImplementedError("annchor_selective_subset needs the reference's materialised per-pair state; "
                                  "use alpha_rss(y, dne, alpha), which evaluates the metric on the device")

    def query(self, Q, nn=15, p_work=0.3, get_exact_query_ijs=None):
        """Query new data against the fitted index (annchor/annchor.py:643-683): returns
        (ngi int64 (len(Q), nn), ngd float64 (len(Q), nn)), the nn approximate nearest points of X per
        query.  ``p_work`` is the fraction of the len(Q) * len(X) brute-force evaluations to spend."""
        if get_exact_query_ijs is not None:
            raise NotImplementedError("annchor_b200 evaluates the metric with its CUDA kernels; a host "
                                      "get_exact_query_ijs would be a CPU fallback")
        if not hasattr(self, "neighbor_graph"):
            raise RuntimeError("query() needs a fitted index: call fit() first")
        nq = len(Q)
        na = self.n_anchors * nq
        nbf = nq * self.nx
        limit = ((nq * nn * 3) // 2 - 1 + na) / nbf
        if p_work < limit:
            print("Warning: p_work too low")
            print("Increasing p_work to %5.3f" % limit)
            p_work = limit
        o = self._order
        if isinstance(self.X, np.ndarray) and self.X.dtype.kind not in "US":
            Xo = np.asarray(self.X) if o is None else np.asarray(self.X)[o]
            both = np.concatenate([Xo, np.asarray(Q, dtype=self.X.dtype)])
        else:
            both = (list(self.X) if o is None else [self.X[i] for i in o]) + list(Q)
        ds = Dataset(self.ctx, both, self.f, cost_matrix=self._cost)
        try:
            ngi, ngd, evals = self._index.query(ds, nq, nn, p_work)
        finally:
            ds.close()
        if o is not None:
            ngi = np.where(ngi >= 0, o[np.maximum(ngi, 0)], -1)
        self.query_evals = evals
        return ngi, ngd

    def fit(self):
        """Computes the approx nearest neighbour graph (annchor/annchor.py:532-623)."""
        origin = time.time()

        def stage(name, fn, *a, **k):
            t = time.time()
            r = fn(*a, **k)
            self.ctx.sync()
            dt = time.time() - t
            self.stage_times[name] = self.stage_times.get(name, 0.0) + dt
            if self.verbose:
                print("%40s: %6.3f | %6.3f" % (name, dt, time.time() - origin))
            return r

        stage("get_anchors", self.get_anchors)
        stage("get_locality", self.get_locality)
        niters = self.niters
        for it in range(niters):
            self._it = it
            try:
                stage("get_sample", self.get_sample)
            except NothingToSample as err:
                if it == 0:
                    raise ValueError("Sampler raised NothingToSample on first iteration.") from err
                print("Warning: main loop terminated early with nothing left to sample.")
                break
            stage("fit_predict_regression", self.fit_predict_regression)
            stage("fit_predict_errors", self.fit_predict_errors)
            stage("select_refine_candidate_pairs", self.select_refine_candidate_pairs, w=1 / niters, it=it)
            if it < niters - 1:
                stage("update_anchor_points", self.update_anchor_points)
        stage("get_ann", self.get_ann)
        return self


class BruteForce:
    """Exact k-NN graph on the B200 -- the reference's ``BruteForce`` (annchor/annchor.py:943-1023):
    ``BruteForce(X, func).fit().neighbor_graph``.

    The reference materialises the full (nx, nx) distance matrix and returns its row-wise argsort /
    sort.  Here ``n_neighbors`` columns are returned (column 0 = the point itself at distance 0, like
    the rows of the reference's graph); ``n_neighbors=None`` means all ``nx`` columns, which is only
    accepted for nx <= 4096.  Euclidean / cosine rows of up to 128 dimensions go through the
    tensor-core path (bf16-split GEMM prunes, the exact metric kernel re-ranks, a per-row
    certificate guarantees exactness; csrc/bruteforce.cu); everything else is all-pairs through the
    metric kernels.  Ties are ordered by neighbour id.
    """

    def __init__(self, X, func, func_kwargs=None, verbose=False, get_exact_ijs=None, backend="loky",
                 n_neighbors=None, device=0, ctx=None, _dataset=None):
        if not isinstance(func, str) or func not in _lib.METRIC_IDS:
            raise NotImplementedError("annchor_b200.BruteForce evaluates one of %s on the GPU; there is no CPU "
                                      "fallback for callables" % sorted(_lib.METRIC_IDS))
        if get_exact_ijs is not None and not isinstance(get_exact_ijs, GpuExactIJs):
            raise NotImplementedError("get_exact_ijs must be None or an annchor_b200.GpuExactIJs")
        cost = (func_kwargs or {}).get("cost_matrix") if func == "wasserstein" else None
        self.X = X
        self.nx = len(X)
        self.f = func
        self.verbose = verbose
        if n_neighbors is None:
            if self.nx > 4096:
                raise ValueError("BruteForce on %d points: pass n_neighbors (the full (nx, nx) sorted matrix "
                                 "the reference returns is not materialised beyond nx = 4096)" % self.nx)
            n_neighbors = self.nx
        self.n_neighbors = int(n_neighbors)
        self.ctx = ctx or default_context(device)
        self._dataset = _dataset if _dataset is not None else Dataset(self.ctx, X, func, cost_matrix=cost)

    def fit(self):
        self.neighbor_graph = self._dataset.bruteforce_knn(self.n_neighbors)
        return self


def compare_neighbor_graphs(nng_1, nng_2, n_neighbors):
    """Tie-aware number of incorrect NN pairs (annchor/annchor.py:1026-1066): per row the multiset
    difference of the first n_neighbors distances rounded to 3 decimals."""
    err = 0
    for ix in range(nng_1[0].shape[0]):
        a = Counter(np.round(nng_1[1][ix][:n_neighbors], 3).astype(np.float32))
        b = Counter(np.round(nng_2[1][ix][:n_neighbors], 3).astype(np.float32))
        err += len(a - b)
    return int(err)
