"""Oracle pipeline: numpy/C restatement of ``Annchor.fit()`` (annchor/annchor.py:532-623).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  The stage methods keep the
reference's order and arithmetic but not its data structures: the per-point pair
index ``I`` (a numba typed Dict in the reference, annchor/utils.py:502-540) is a
CSR pair (row_ptr, row_pairs) here, and the Python row loops are C loops.

Known, documented deviations from the reference (none changes a value on the
inputs the reference's tests use):
  * ``update_anchor_points`` has no 10 s wall-clock cut-off (annchor/annchor.py:511);
    all chunks are processed.
  * the CSR is the intended one; the reference's end-of-array bookkeeping quirk
    (annchor/utils.py:516-521) is not reproduced.
  * sort ties (np.argsort quicksort) are resolved by index where the reference
    leaves them unspecified.
"""
from collections import Counter

import numpy as np

from .clib import lib, ptr
from .metrics import PairMetric

FEATURE_NAMES = ["lower bound", "upper bound", "double anchor distance", "is anchor"]


class NothingToSample(Exception):
    pass


# ---------------------------------------------------------------------------
# leaves
# ---------------------------------------------------------------------------

def get_bounds_ijs(IJs, D):
    """annchor/utils.py:274-301"""
    IJs = np.ascontiguousarray(IJs, dtype=np.int64)
    D = np.ascontiguousarray(D, dtype=np.float64)
    out = np.empty((IJs.shape[0], 2))
    lib().orc_bounds_ijs(ptr(IJs), IJs.shape[0], ptr(D), D.shape[1], ptr(out))
    return out


def get_dad_ijs(IJs, D):
    """annchor/utils.py:355-380"""
    IJs = np.ascontiguousarray(IJs, dtype=np.int64)
    D = np.ascontiguousarray(D, dtype=np.float64)
    out = np.empty(IJs.shape[0])
    lib().orc_dad_ijs(ptr(IJs), IJs.shape[0], ptr(D), D.shape[0], D.shape[1], ptr(out))
    return out


def row_kth(RA, row_ptr, row_pairs, k):
    """thresh loop, annchor/annchor.py:399-404"""
    nx = row_ptr.shape[0] - 1
    out = np.empty(nx)
    lib().orc_row_kth(ptr(RA), ptr(row_ptr), ptr(row_pairs), nx, k, ptr(out))
    return out


def get_probs(p, labels, errs):
    """annchor/utils.py:581-589; errs = list of sorted float64 arrays."""
    eptr = np.zeros(len(errs) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in errs], out=eptr[1:])
    flat = np.ascontiguousarray(np.concatenate(errs), dtype=np.float64)
    p = np.ascontiguousarray(p, dtype=np.float64)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    out = np.empty(p.shape[0])
    lib().orc_probs(ptr(p), ptr(labels), p.shape[0], ptr(flat), ptr(eptr), ptr(out))
    return out


def get_nn(nx, nn, RA, IJs, row_ptr, row_pairs, ncm):
    """annchor/utils.py:383-429"""
    ngi = np.zeros((nx, nn - 1), dtype=np.int64)
    ngd = np.zeros((nx, nn - 1))
    m8 = np.ascontiguousarray(ncm, dtype=np.uint8)
    lib().orc_get_nn(nx, nn, ptr(RA), ptr(IJs), ptr(row_ptr), ptr(row_pairs), ptr(m8),
                     ptr(ngi), ptr(ngd))
    return ngi, ngd


def update_bounds(IJs, kptr, kids, kds):
    """annchor/utils.py:326-352 over CSR known-distance lists sorted by id."""
    IJs = np.ascontiguousarray(IJs, dtype=np.int64)
    out = np.empty((IJs.shape[0], 2))
    lib().orc_update_bounds(ptr(IJs), IJs.shape[0], ptr(kptr), ptr(kids), ptr(kds), ptr(out))
    return out


class NumbaRNG:
    """numba's per-thread MT19937 as driven by annchor/utils.py:572,555-557."""

    def __init__(self, seed):
        L = lib()
        self._st = np.zeros(L.orc_mt_state_size() // 4 + 1, dtype=np.uint32)
        L.orc_mt_seed(ptr(self._st), int(seed) & 0xFFFFFFFF)

    def choice_no_replace(self, a, size):
        x = np.ascontiguousarray(a, dtype=np.int64).copy()
        lib().orc_numba_shuffle(ptr(self._st), ptr(x), x.shape[0])
        return x[:size]


def stratified_partition(sample_feature, n_samples, n_partitions=7):
    """SimpleStratifiedSampler.get_partition, annchor/samplers.py:119-140"""
    n = sample_feature.shape[0]
    iq1, iq3 = int(n / 100), int(99 * n / 100)
    if iq1 * n_partitions < n_samples:
        iq1, iq3 = int(n / 10), int(9 * n / 10)
    if iq1 * n_partitions < n_samples:
        n_samples = iq1 * n_partitions
    q1 = np.partition(sample_feature, iq1)[iq1]
    q3 = np.partition(sample_feature, iq3)[iq3]
    bins = np.hstack([-np.inf, np.linspace(q1, q3, n_partitions - 1), np.inf])
    return bins, n_samples


def fit_stratified_linear(F, X3, y, bins):
    """SimpleStratifiedLinearRegression.fit, annchor/regressors.py:39-67.
    sklearn's LinearRegression(fit_intercept=True) == least squares on centred
    data; returns coef (n_bins,3) and intercept (n_bins,)."""
    nb = bins.shape[0] - 1
    coef = np.zeros((nb, X3.shape[1]))
    icpt = np.zeros(nb)
    for b in range(nb):
        m = (F > bins[b]) & (F <= bins[b + 1])
        Xb, yb = X3[m], y[m]
        if Xb.shape[0] == 0:
            raise ValueError("empty regression bin %d" % b)
        xm, ym = Xb.mean(axis=0), yb.mean()
        c = np.linalg.lstsq(Xb - xm, yb - ym, rcond=None)[0]
        coef[b] = c
        icpt[b] = ym - xm @ c
    return coef, icpt


def predict_stratified_linear(F, X3, bins, coef, icpt):
    """SimpleStratifiedLinearRegression.predict, annchor/regressors.py:71-103"""
    y = np.zeros(X3.shape[0])
    for b in range(bins.shape[0] - 1):
        m = (F > bins[b]) & (F <= bins[b + 1])
        if m.any():
            y[m] = X3[m] @ coef[b] + icpt[b]
    return y


def error_labels(F, bins):
    """SimpleStratifiedErrorRegression.predict, annchor/error_predictors.py:56-67
    (closed intervals, later bins overwrite)."""
    labels = np.empty(F.shape[0], dtype=np.int64)
    for b in range(bins.shape[0] - 1):
        labels[(F >= bins[b]) & (F <= bins[b + 1])] = b
    return labels


def compare_neighbor_graphs(nng_1, nng_2, n_neighbors):
    """annchor/annchor.py:1026-1066: tie-aware count of wrong NN distances."""
    err = 0
    for ix in range(nng_1[0].shape[0]):
        a = Counter(np.round(nng_1[1][ix][:n_neighbors], 3).astype(np.float32))
        b = Counter(np.round(nng_2[1][ix][:n_neighbors], 3).astype(np.float32))
        err += len(a - b)
    return int(err)


# ---------------------------------------------------------------------------
# pipeline
# ---------------------------------------------------------------------------

class OracleAnnchor:
    """Restatement of ``annchor.Annchor`` restricted to ``fit()`` with the
    default plug-ins (MaxMin picker, SimpleStratified sampler / regression /
    error predictor).  ``metric`` is a name understood by
    ``oracle.metrics.PairMetric`` or a callable ``IJ -> float64[len(IJ)]``."""

    def __init__(self, X, metric, n_anchors=20, n_neighbors=15, n_samples=5000, p_work=0.1,
                 random_seed=42, locality=5, loc_thresh=1, loc_min=None, is_metric=True,
                 niters=2, lookahead=5, anchors=None, trace=None):
        self.X = X
        self.nx = len(X)
        self.N = (self.nx * (self.nx - 1)) // 2
        self.pair_fn = PairMetric(X, metric) if isinstance(metric, str) else metric
        self.evals = 0
        self.n_anchors = n_anchors
        # annchor/annchor.py:126
        self.na = int(np.sum([self.nx - j for j in range(1, n_anchors + 1)]))
        self.n_neighbors = n_neighbors
        self.n_samples = n_samples
        # p_work clamps, annchor/annchor.py:132-142
        self.p_work = min(p_work, 1.0)
        min_p_work = min((2 * (self.na + self.n_samples) + 1) / self.N, 1)
        if self.p_work < min_p_work:
            self.p_work = min_p_work
        self.random_seed = random_seed
        self.locality = locality
        self.loc_thresh = loc_thresh
        self.loc_min = 10 * n_neighbors if loc_min is None else loc_min
        self.loc_min = int(np.clip(self.loc_min, 0, self.nx - 1))
        self.is_metric = is_metric
        self.niters = niters
        self.lookahead = lookahead
        self.fixed_anchors = anchors
        self.RefineApprox = None
        self.loop_num = 0
        self.trace = trace  # optional dict collecting per-stage arrays

    def _t(self, key, val):
        if self.trace is not None:
            self.trace[key] = np.array(val, copy=True)

    # -- stage 1: anchors (annchor/pickers.py:18-52, 86-107) ------------------
    def get_anchors(self):
        nx, na = self.nx, self.n_anchors
        D = np.full((na, nx), np.inf)
        cols = np.arange(nx, dtype=np.int64)
        if self.fixed_anchors is not None:
            A = np.asarray(self.fixed_anchors, dtype=np.int64)
            for i, a in enumerate(A):
                D[i] = self.pair_fn(np.stack([np.full(nx, a, dtype=np.int64), cols], axis=1))
        else:
            rs = np.random.RandomState(self.random_seed)
            A = np.zeros(na, dtype=np.int64)
            ix = int(rs.randint(nx))
            for i in range(na):
                A[i] = ix
                D[i] = self.pair_fn(np.stack([np.full(nx, ix, dtype=np.int64), cols], axis=1))
                # anchor 0 is excluded from the min after round 0 (pickers.py:47-50)
                ix = int(np.argmax(D[0] if i == 0 else D[1:i + 1].min(axis=0)))
        self.A = A
        self.D = np.ascontiguousarray(D.T)
        self.evals += na * nx

    # -- candidate set (annchor/annchor.py:208-256, utils.py:437-540) ---------
    def get_locality(self):
        nx, na, L = self.nx, self.n_anchors, self.locality
        self.sid = np.argsort(self.D, axis=1, kind="stable")[:, :L]
        M = np.zeros((nx, na), dtype=np.int32)
        np.put_along_axis(M, self.sid, 1, axis=1)
        loc_min = min(self.loc_min, nx - 1)
        t = np.empty(nx, dtype=np.int64)
        step = max(1, (1 << 24) // nx)
        for s in range(0, nx, step):
            cnt = M[s:s + step] @ M.T
            kth = -np.partition(-cnt, loc_min, axis=1)[:, loc_min]
            t[s:s + step] = np.minimum(kth, self.loc_thresh)
        self.loc_t = t
        chunks = []
        for s in range(0, nx, step):
            cnt = M[s:s + step] @ M.T
            thr = np.minimum(t[s:s + step, None], t[None, :])
            ii, jj = np.nonzero(cnt >= thr)
            ii += s
            keep = jj > ii
            chunks.append(np.stack([ii[keep], jj[keep]], axis=1))
        self.IJs = np.ascontiguousarray(np.concatenate(chunks), dtype=np.int64)
        P = self.IJs.shape[0]
        # CSR: pairs where the point is the 2nd index (ascending), then 1st index
        ids = np.arange(P, dtype=np.int64)
        o2 = np.argsort(self.IJs[:, 1], kind="stable")
        c1 = np.bincount(self.IJs[:, 0], minlength=nx)
        c2 = np.bincount(self.IJs[:, 1], minlength=nx)
        self.row_ptr = np.zeros(nx + 1, dtype=np.int64)
        np.cumsum(c1 + c2, out=self.row_ptr[1:])
        self.row_pairs = np.empty(2 * P, dtype=np.int64)
        s2 = np.zeros(nx + 1, dtype=np.int64)
        np.cumsum(c2, out=s2[1:])
        s1 = np.zeros(nx + 1, dtype=np.int64)
        np.cumsum(c1, out=s1[1:])
        for_second = self.row_ptr[:-1]
        pos2 = np.repeat(for_second, c2) + (np.arange(P) - np.repeat(s2[:-1], c2))
        self.row_pairs[pos2] = o2
        pos1 = np.repeat(for_second + c2, c1) + (np.arange(P) - np.repeat(s1[:-1], c1))
        self.row_pairs[pos1] = ids
        if np.any(c1 + c2 < self.n_neighbors):
            raise Exception("Error: Not enough candidates in pool for all indices.\n"
                            "Try again with higher locality.")

    def row(self, i):
        return self.row_pairs[self.row_ptr[i]:self.row_ptr[i + 1]]

    # -- stage 2: features (annchor/annchor.py:258-311) -----------------------
    def get_features(self):
        P = self.IJs.shape[0]
        dad = get_dad_ijs(self.IJs, self.D)
        bounds = get_bounds_ijs(self.IJs, self.D)
        anchors = np.zeros(P)
        for a in self.A:
            anchors[self.row(a)] = 1
        self.features = np.ascontiguousarray(np.vstack([bounds.T, dad, anchors]).T)
        self.not_computed_mask = self.features[:, 3] < 1

    # -- sampler (annchor/samplers.py:75-140, utils.py:543-578) ---------------
    def get_sample(self):
        ncm = self.not_computed_mask
        if not ncm.any():
            raise NothingToSample()
        sf = self.features[ncm][:, 2]
        indices = np.arange(ncm.shape[0])[ncm]
        bins, n_samples = stratified_partition(sf, self.n_samples)
        if n_samples == 0:
            raise NothingToSample()
        nb = bins.shape[0] - 1
        bin_size, rem = n_samples // nb, n_samples % nb
        rng = NumbaRNG(self.random_seed + self.loop_num)
        self.loop_num += 1
        parts = []
        for b in range(nb):
            ixmask = indices[(sf >= bins[b]) & (sf < bins[b + 1])]
            want = bin_size + (b < rem)
            parts.append(ixmask if ixmask.shape[0] < want else rng.choice_no_replace(ixmask, want))
            if parts[-1].shape[0] < 2:
                raise Exception("Some sampler bins contain too few samples")
        self.sample_ixs = np.hstack(parts)
        self.n_samples = self.sample_ixs.shape[0]
        self.sample_bins = bins
        self.sample_features = self.features[self.sample_ixs]
        self.sample_y = self.pair_fn(self.IJs[self.sample_ixs])
        self.not_computed_mask[self.sample_ixs] = False
        self.evals += self.sample_y.shape[0]

    # -- stage 3 (annchor/annchor.py:345-393) ---------------------------------
    def fit_predict_regression(self):
        F = self.sample_features[:, 2]
        self.coef, self.icpt = fit_stratified_linear(
            F, self.sample_features[:, :3], self.sample_y, self.sample_bins)
        pred = predict_stratified_linear(
            self.features[:, 2], self.features[:, :3], self.sample_bins, self.coef, self.icpt)
        self.sample_predict = pred[self.sample_ixs]
        pred = np.clip(pred, self.features[:, 0], self.features[:, 1])
        if not self.is_metric:
            for i, a in enumerate(self.A):
                ra = self.row(a)
                ijs = self.IJs[ra]
                other = np.sum(ijs * (ijs != a), axis=1)
                pred[ra] = self.D[other, i]
        self.pred = pred
        if self.RefineApprox is None:
            self.RefineApprox = pred.copy()
        else:
            m = self.not_computed_mask
            self.RefineApprox[m] = pred[m]
        self.RefineApprox[self.sample_ixs] = self.sample_y

    def fit_predict_errors(self):
        F = self.sample_features[:, 2]
        err = self.sample_y - self.sample_predict
        bins = self.sample_bins
        self.errs = [np.sort(err[(F >= bins[b]) & (F <= bins[b + 1])])
                     for b in range(bins.shape[0] - 1)]
        self.errors = error_labels(self.features[:, 2], bins)

    # -- stage 4 (annchor/annchor.py:395-473, utils.py:606-621) ---------------
    def guarantee_nmin(self, nmin):
        RA, ncm = self.RefineApprox, self.not_computed_mask
        for i in range(self.nx):
            ri = self.row(i)
            m = ncm[ri]
            n_todo = nmin - int(np.sum(~m))
            if n_todo > 0:
                vals = RA[ri][m]
                kth = np.partition(vals, n_todo)[n_todo]
                RA[ri[m][vals < kth]] = -1

    def select_refine_candidate_pairs(self, w, it):
        nn = self.n_neighbors
        self.thresh = row_kth(self.RefineApprox, self.row_ptr, self.row_pairs, nn)
        if it == 0:
            self.guarantee_nmin(3 * nn // 2)
        ncm = self.not_computed_mask
        RA = self.RefineApprox
        p = np.maximum(self.thresh[self.IJs[:, 0]] - RA, self.thresh[self.IJs[:, 1]] - RA)[ncm]
        prob = get_probs(p, self.errors[ncm], self.errs)
        self.prob = prob
        n_refine = int((self.p_work * self.N - self.na - self.n_samples) * w) + 1
        n_refine = max(n_refine, 0)
        self.n_refine = n_refine
        if n_refine >= prob.shape[0]:
            cand = np.arange(prob.shape[0])
            nxt = np.arange(prob.shape[0])
        else:
            if n_refine * self.lookahead >= prob.shape[0]:
                large = np.arange(prob.shape[0])
            else:
                large = np.argpartition(-prob, n_refine * self.lookahead)[:n_refine * self.lookahead]
            ap = np.argpartition(-prob[large], n_refine)
            cand = large[ap[:n_refine]]
            nxt = large[ap[n_refine:]]
        back = np.arange(ncm.shape[0])[ncm]
        self.nextback = back[nxt]
        mapback = back[cand]
        self.mapback = mapback
        exact = self.pair_fn(self.IJs[mapback])
        self.evals += exact.shape[0]
        RA[mapback] = exact
        ncm[mapback] = False

    # -- stage 2b (annchor/annchor.py:475-512) --------------------------------
    def known_lists(self):
        """Per-point (ids, dists) of computed pairs, ids ascending, as CSR."""
        known = np.nonzero(~self.not_computed_mask)[0]
        ij = self.IJs[known]
        d = self.RefineApprox[known]
        src = np.concatenate([ij[:, 0], ij[:, 1]])
        dst = np.concatenate([ij[:, 1], ij[:, 0]])
        dd = np.concatenate([d, d])
        o = np.lexsort((dst, src))
        kptr = np.zeros(self.nx + 1, dtype=np.int64)
        np.cumsum(np.bincount(src, minlength=self.nx), out=kptr[1:])
        return kptr, np.ascontiguousarray(dst[o]), np.ascontiguousarray(dd[o])

    def update_anchor_points(self):
        mapback = self.nextback
        if mapback.shape[0] == 0:
            return
        kptr, kids, kds = self.known_lists()
        b = update_bounds(self.IJs[mapback], kptr, kids, kds)
        self.features[mapback, 0] = np.maximum(b[:, 0], self.features[mapback, 0])
        self.features[mapback, 1] = np.minimum(b[:, 1], self.features[mapback, 1])

    # -- stage 5 (annchor/annchor.py:514-530) ---------------------------------
    def get_ann(self):
        ngi, ngd = get_nn(self.nx, self.n_neighbors, self.RefineApprox, self.IJs,
                          self.row_ptr, self.row_pairs, self.not_computed_mask)
        self.neighbor_graph = (
            np.hstack([np.arange(self.nx)[:, None], ngi]),
            np.hstack([np.zeros((self.nx, 1)), ngd]),
        )

    def fit(self):
        self.get_anchors()
        self._t("A", self.A)
        self._t("D", self.D)
        self.get_locality()
        self._t("IJs", self.IJs)
        self.get_features()
        self._t("features0", self.features)
        for it in range(self.niters):
            try:
                self.get_sample()
            except NothingToSample as err:
                if it == 0:
                    raise ValueError("Sampler raised NothingToSample on first iteration.") from err
                break
            self._t("sample_ixs%d" % it, self.sample_ixs)
            self._t("sample_bins%d" % it, self.sample_bins)
            self.fit_predict_regression()
            self._t("pred%d" % it, self.pred)
            self.fit_predict_errors()
            self._t("errors%d" % it, self.errors)
            self.select_refine_candidate_pairs(1 / self.niters, it)
            self._t("thresh%d" % it, self.thresh)
            self._t("prob%d" % it, self.prob)
            self._t("mapback%d" % it, np.sort(self.mapback))
            if it < self.niters - 1:
                self.update_anchor_points()
                self._t("features_upd%d" % it, self.features)
        self.get_ann()
        return self


class OracleBruteForce:
    """annchor/annchor.py:943-1023 (all pairs + full argsort)."""

    def __init__(self, X, metric):
        self.nx = len(X)
        self.pair_fn = PairMetric(X, metric) if isinstance(metric, str) else metric

    def fit(self):
        nx = self.nx
        iu = np.triu_indices(nx, 1)
        d = self.pair_fn(np.stack(iu, axis=1))
        D = np.zeros((nx, nx))
        D[iu] = d
        D = D + D.T
        self.D = D
        self.neighbor_graph = (np.argsort(D, axis=1, kind="stable"), np.sort(D, axis=1))
        return self
