"""Host-side plug-ins with the reference's duck-typed protocols (annchor/pickers.py,
samplers.py, regressors.py, error_predictors.py), re-designed for the streaming index: the
Theta(N^2) feature arrays the reference hands to its plug-ins do not exist here, so

  * anchor pickers keep the reference protocol exactly: ``get_anchors(ann) -> (A, D, n_evals)``;
  * the sampler protocol is ``sample_index(ann) -> (sample_ijs, sample_features, sample_bins)``
    (device reductions + host choice) instead of ``sample(features, ...)``;
  * regression / error predictor keep ``fit(sample_features, feature_names, y, sample_bins)`` and
    must expose what the device needs to predict every pair: bin edges + per-bin linear
    coefficients, and per-bin sorted error tables.  The reference's own
    SimpleStratifiedLinearRegression / SimpleStratifiedErrorRegression instances satisfy this.
"""
import numpy as np

FEATURE_NAMES = ["lower bound", "upper bound", "double anchor distance", "is anchor"]


class NothingToSample(Exception):
    """annchor/samplers.py:18"""


# ---------------------------------------------------------------------------------------------
# anchor pickers (annchor/pickers.py)
# ---------------------------------------------------------------------------------------------
class MaxMinAnchorPicker:
    """annchor/pickers.py:18-52.  All n_anchors rounds run on the device (annb_index_maxmin);
    only the first anchor comes from the host RNG, drawn exactly as the reference draws it."""

    device_native = True

    def get_anchors(self, ann):
        first = int(np.random.RandomState(ann.random_seed).randint(ann.nx))
        A = ann._index.maxmin(first)
        return A, None, ann.n_anchors * ann.nx


class ExternalAnchorPicker:
    """annchor/pickers.py:55-83: the anchors are items that need not belong to X.  Returns an empty
    ``A`` like the reference, so no pair is pre-marked as computed.  The distances X -> anchors are
    evaluated on the device over a temporary data set holding X followed by the anchor items."""

    def __init__(self, A):
        self.A = A
        self.is_anchor_safe = False

    def get_anchors(self, ann):
        from .core import Dataset
        nx, na = ann.nx, ann.n_anchors
        items = list(self.A)[:na]
        if len(items) != na:
            raise ValueError("ExternalAnchorPicker needs n_anchors=%d items, got %d" % (na, len(items)))
        if isinstance(ann.X, np.ndarray) and ann.X.dtype.kind not in "US":
            both = np.concatenate([np.asarray(ann.X), np.asarray(items, dtype=ann.X.dtype).reshape(na, -1)])
        else:
            both = list(ann.X) + items
        tmp = Dataset(ann.ctx, both, ann.f, cost_matrix=ann._cost)
        try:
            D = tmp.anchor_dists(np.arange(nx, nx + na, dtype=np.int64))[:nx]
        finally:
            tmp.close()
        return np.array([], dtype=np.int64), np.ascontiguousarray(D), na * nx


class SelectedAnchorPicker:
    """annchor/pickers.py:86-107: anchors given as indices into X."""

    def __init__(self, A):
        self.A = np.asarray(A, dtype=np.int64)

    def get_anchors(self, ann):
        return self.A, ann._dataset.anchor_dists(self.A), ann.n_anchors * ann.nx


class RandomAnchorPicker:
    """annchor/pickers.py:110-128"""

    def get_anchors(self, ann):
        rs = np.random.RandomState(ann.random_seed)
        A = rs.choice(np.arange(ann.nx), ann.n_anchors, replace=False).astype(np.int64)
        return A, ann._dataset.anchor_dists(A), ann.n_anchors * ann.nx


# ---------------------------------------------------------------------------------------------
# sampler (annchor/samplers.py:113-140 + utils.py:543-578)
# ---------------------------------------------------------------------------------------------
class NumbaRNG:
    """numba's in-@njit np.random stream (MT19937 + Fisher-Yates) as used by the reference's
    sampler (annchor/utils.py:555-557,572), via the host helper in libannb."""

    def __init__(self, seed):
        import ctypes as C
        from . import _lib
        self._L = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._L.annb_numba_rng_new(int(seed) & 0xFFFFFFFF, C.byref(self._h)))

    def choice_no_replace(self, a, size):
        from . import _lib
        x = np.ascontiguousarray(a, dtype=np.int64).copy()
        _lib.check(self._L.annb_numba_rng_shuffle(self._h, _lib.ptr(x), x.shape[0]))
        return x[:size]

    def __del__(self):
        try:
            self._L.annb_numba_rng_free(self._h)
        except Exception:
            pass


def bin_of(v, inner):
    """Sampler bin of every value: bins[b] <= v < bins[b + 1] (annchor/utils.py:547-549) = the number of inner edges
    <= v -- what np.digitize(v, inner) returns, without its per-element binary search (int8: at most 8 bins)."""
    out = np.zeros(v.shape[0], dtype=np.int8)
    for e in inner:
        out += v >= e
    return out


def host_features(D, ijs):
    """[lb, ub, dad] of explicit pairs in float64 from the host copy of D (annchor/utils.py:274-301,
    355-380) -- used for the (<= n_samples) sample pairs only."""
    I, J = ijs[:, 0], ijs[:, 1]
    Di, Dj = D[I], D[J]
    lb = np.abs(Di - Dj).max(axis=1)
    ub = (Di + Dj).min(axis=1)
    cA = np.argmin(D, axis=1)
    dad = (D[I, cA[J]] + D[J, cA[I]]) / 2
    return np.stack([lb, ub, dad], axis=1)


class SimpleStratifiedSampler:
    """Stratified sample over the not-computed candidate pairs, binned by double anchor distance
    (annchor/samplers.py:113-140, annchor/utils.py:543-578).

    The device returns a uniform pool of the not-computed candidates (annb_index_sample_pool).
      * exact mode -- at most ``exact_limit`` such pairs exist (N up to a few thousand): the pool is
        all of them; sorted by (i, j) it is the reference's ``indices`` array, dad is recomputed in
        float64 from D, and the reference's algorithm runs verbatim including numba's MT19937
        Fisher-Yates draw -- the sample equals the reference's bit for bit.
      * pool mode -- larger problems: the 1 % / 99 % order statistics and the per-bin uniform
        draws are taken on the hash-selected pool (a uniform sub-sample of ~pool_size pairs), since
        the reference's draw needs the materialised Theta(N^2) pair list."""

    def __init__(self, partition_feature_name="double anchor distance", n_partitions=7,
                 exact_limit=4_000_000, pool_size=1 << 18):
        self.partition_feature_name = partition_feature_name
        self.n_partitions = n_partitions
        self.exact_limit = exact_limit   # at most this many not-computed candidates -> exact mode
        self.pool_size = pool_size       # size of the uniform pool otherwise
        self.loop_num = 0

    def get_partition(self, sample_feature, n_samples, n_total=None):
        """annchor/samplers.py:119-140; n_total = size of the full population when sample_feature
        is only a pool of it."""
        m = sample_feature.shape[0]
        n = m if n_total is None else n_total
        iq1, iq3, tenth = int(n / 100), int(99 * n / 100), False
        if (iq1 * self.n_partitions) < n_samples:
            iq1, iq3, tenth = int(n / 10), int(9 * n / 10), True
        if (iq1 * self.n_partitions) < n_samples:
            n_samples = iq1 * self.n_partitions
            print("Warning: n_samples too large for data set size.\n"
                  + "Reducing n_samples to %d." % n_samples)
        if n_total is not None:  # the order statistics are taken on the pool
            iq1, iq3 = (int(m / 10), int(9 * m / 10)) if tenth else (int(m / 100), int(99 * m / 100))
        q1 = np.partition(sample_feature, iq1)[iq1]
        q3 = np.partition(sample_feature, iq3)[iq3]
        bins = np.hstack([-np.inf, np.linspace(q1, q3, self.n_partitions - 1), np.inf])
        return bins, n_samples

    def sample_index(self, ann):
        ix = ann._index
        seed = int(ann.random_seed) + self.loop_num
        n_nc = ix.stats()["n_not_computed"]
        n_pool, n_nc, exact = ix.sample_pool(seed, self.exact_limit if n_nc <= self.exact_limit
                                             else self.pool_size)
        if n_nc <= 0 or n_pool == 0:
            raise NothingToSample()
        ijs, dad = ix.get_pool()
        if exact:
            order = np.argsort(ijs[:, 0] * np.int64(ann.nx) + ijs[:, 1], kind="stable")
            ijs = ijs[order]
            D = ann.D
            cA = np.argmin(D, axis=1)
            sf = (D[ijs[:, 0], cA[ijs[:, 1]]] + D[ijs[:, 1], cA[ijs[:, 0]]]) / 2  # float64, utils.py:378-380
            bins, n_samples = self.get_partition(sf, ann.n_samples)
        else:
            sf = dad
            bins, n_samples = self.get_partition(sf, ann.n_samples, n_total=n_nc)
        if n_samples != ann.n_samples:
            print("Warning: n_samples has changed from %d to %d." % (ann.n_samples, n_samples))
        if n_samples == 0:
            raise NothingToSample()
        bin_size, rem = n_samples // self.n_partitions, n_samples % self.n_partitions
        P = self.n_partitions
        inner = bins[1:-1]
        rng = NumbaRNG(seed) if exact else None
        bidx = bin_of(sf, inner)

        def priorities(ijs):
            # order-independent priorities (the pool arrives in atomic order): splitmix64 of the pair
            key = (ijs[:, 0].astype(np.uint64) << np.uint64(32)) | ijs[:, 1].astype(np.uint64)
            with np.errstate(over="ignore"):
                z = key + np.uint64((seed * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
                z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
                z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
                return z ^ (z >> np.uint64(31))

        if not exact:
            # bins the uniform pool left short are re-sampled at their own rate over all tiles (the
            # reference draws per bin from the materialised pair list, so a rare dad range is as well
            # represented as a common one)
            want_b = n_samples // P + 1
            cnt = np.bincount(bidx, minlength=P)
            short = cnt < want_b
            if short.any():
                frac = max(n_pool / float(n_nc), 1e-12)
                rate = np.where(short, np.minimum(1.0, 8.0 * want_b * frac / np.maximum(cnt, 0.25)), 0.0)
                n2 = ix.sample_pool_bins(seed, bins, rate, self.pool_size)
                if n2 > 0:
                    ijs2, dad2 = ix.get_pool()
                    keep = ~short[bidx]
                    ijs = np.concatenate([ijs[keep], ijs2])
                    sf = np.concatenate([sf[keep], dad2])
                    bidx = np.concatenate([bidx[keep], bin_of(dad2, inner)])
            prio = priorities(ijs)
        self.loop_num += 1
        parts = []
        for b in range(P):
            ixmask = np.flatnonzero(bidx == b)  # members of the bin in increasing pool index
            want = bin_size + (b < rem)
            if ixmask.shape[0] < want:
                parts.append(ixmask)
            elif exact:
                parts.append(rng.choice_no_replace(ixmask, want))
            else:  # uniform without replacement: the `want` smallest priorities of the bin
                parts.append(ixmask[np.argpartition(prio[ixmask], want - 1)[:want]] if want > 0 else ixmask[:0])
            if parts[-1].shape[0] < 2:
                raise Exception("Some sampler bins contain too few samples")
        sel = np.hstack(parts)
        if n_samples != sel.shape[0]:
            print("Warning: Some bins contained fewer samples than requested")
        sample_ijs = np.ascontiguousarray(ijs[sel])
        if exact:
            # float64 anchor bounds, overlaid with the bounds update_anchor_points has tightened:
            # the reference trains on features[sample_ixs], which carry them (annchor.py:338,503-510)
            feats = host_features(ann.D, sample_ijs)
            kind, a, b = ix.pair_state(sample_ijs)
            t = kind == 2
            feats[t, 0] = np.maximum(feats[t, 0], a[t])
            feats[t, 1] = np.minimum(feats[t, 1], b[t])
        else:
            feats = ix.pair_features(sample_ijs)
        return sample_ijs, feats, bins


MATERIALISED_LIMIT = 20_000_000  # not-computed candidate pairs a reference-protocol sampler may see


class MaterialisedSamplerAdapter:
    """Serves a reference-protocol sampler -- ``sample(features, feature_names, n_samples,
    not_computed_mask, random_seed) -> (sample_ixs, n_actual, sample_bins)`` (annchor/samplers.py:75-110;
    the reference's own SimpleStratifiedSampler / ClusterSampler instances qualify) -- from a
    materialised view of the not-computed candidate pairs: rows in the reference's IJs order
    (sorted by (i, j)), columns [lb, ub, dad, is_anchor=0] in the sweeps' float32 arithmetic
    including tightened bounds.  Limited to MATERIALISED_LIMIT pairs (the reference itself needs
    ~200 B of host memory per candidate pair)."""

    def __init__(self, sampler):
        self.sampler = sampler

    def sample_index(self, ann):
        ix = ann._index
        n_nc = ix.stats()["n_not_computed"]
        if n_nc > MATERIALISED_LIMIT:
            raise NotImplementedError(
                "%d not-computed candidate pairs: a sampler with the reference protocol sample(features, ...) "
                "needs them materialised (limit %d).  Use a sample_index(ann) sampler "
                "(annchor_b200.plugins.SimpleStratifiedSampler) at this size." % (n_nc, MATERIALISED_LIMIT))
        n_pool, n_nc, exact = ix.sample_pool(int(ann.random_seed), max(n_nc + 1024, 1024))
        if n_pool == 0:
            raise NothingToSample()
        ijs, _ = ix.get_pool()
        ijs = ijs[np.argsort(ijs[:, 0] * np.int64(ann.nx) + ijs[:, 1], kind="stable")]
        feats = np.zeros((n_pool, 4))
        feats[:, :3] = ix.pair_features(ijs)
        try:
            sample_ixs, _n, bins = self.sampler.sample(feats, list(FEATURE_NAMES), ann.n_samples,
                                                       np.ones(n_pool, dtype=bool), ann.random_seed)
        except Exception as e:  # the reference raises its own NothingToSample class
            if type(e).__name__ == "NothingToSample":
                raise NothingToSample() from e
            raise
        sample_ixs = np.asarray(sample_ixs, dtype=np.int64)
        return np.ascontiguousarray(ijs[sample_ixs]), feats[sample_ixs, :3], np.asarray(bins, dtype=np.float64)


# ---------------------------------------------------------------------------------------------
# regression (annchor/regressors.py:18-103)
# ---------------------------------------------------------------------------------------------
class _Linear:
    def __init__(self):
        self.coef_ = None
        self.intercept_ = 0.0

    def fit(self, X, y):
        # ordinary least squares with intercept (what sklearn's LinearRegression solves)
        if X.shape[0] == 0:
            raise ValueError("Found array with 0 sample(s) while fitting a regression bin")
        xm, ym = X.mean(axis=0), y.mean()
        c = np.linalg.lstsq(X - xm, y - ym, rcond=None)[0]
        self.coef_ = c
        self.intercept_ = float(ym - xm @ c)
        return self

    def predict(self, X):
        return X @ self.coef_ + self.intercept_


class SimpleStratifiedLinearRegression:
    def __init__(self, reg_feature_names=("lower bound", "upper bound", "double anchor distance"),
                 partition_feature_name="double anchor distance", n_partitions=7):
        self.n_partitions = n_partitions
        self.LRs = [_Linear() for _ in range(n_partitions)]
        self.partition_feature_name = partition_feature_name
        self.reg_feature_names = list(reg_feature_names)

    def _cols(self, feature_names):
        ip = feature_names.index(self.partition_feature_name)
        cols = [i for i, nm in enumerate(feature_names) if nm in self.reg_feature_names]
        return ip, cols

    def fit(self, sample_features, feature_names, sample_y, sample_bins=None):
        ip, cols = self._cols(feature_names)
        F = sample_features[:, ip]
        if sample_bins is None:
            n = F.shape[0]
            q1 = np.partition(F, int(n / 100))[int(n / 100)]
            q3 = np.partition(F, int(99 * n / 100))[int(99 * n / 100)]
            self.sample_bins = np.hstack([-np.inf, np.linspace(q1, q3, self.n_partitions - 1), np.inf])
        else:
            self.n_partitions = sample_bins.shape[0] - 1
            self.sample_bins = sample_bins
        while len(self.LRs) < self.n_partitions:
            self.LRs.append(_Linear())
        for b in range(self.n_partitions):
            m = (F > self.sample_bins[b]) & (F <= self.sample_bins[b + 1])
            self.LRs[b].fit(sample_features[m][:, cols], sample_y[m])

    def predict(self, features, feature_names):
        ip, cols = self._cols(feature_names)
        X, F = features[:, cols], features[:, ip]
        y = np.zeros(X.shape[0])
        for b in range(self.n_partitions):
            m = (F > self.sample_bins[b]) & (F <= self.sample_bins[b + 1])
            if m.any():
                y[m] = self.LRs[b].predict(X[m])
        return y


def regression_device_spec(reg, feature_names):
    """(bins, coef (nb,3), icpt (nb,)) of a fitted stratified-linear regression object --
    ours or the reference's (annchor/regressors.py:18-67: .sample_bins, .LRs[b].coef_/.intercept_)."""
    if not (hasattr(reg, "sample_bins") and hasattr(reg, "LRs")):
        raise NotImplementedError(
            "the streaming index can only evaluate regressions that expose `sample_bins` and "
            "per-bin linear models `LRs[b].coef_ / .intercept_` (the reference's "
            "SimpleStratifiedLinearRegression does); arbitrary predict() callables would need the "
            "materialised Theta(N^2) feature array")
    names = list(getattr(reg, "reg_feature_names", FEATURE_NAMES[:3]))
    bins = np.asarray(reg.sample_bins, dtype=np.float64)
    nb = bins.shape[0] - 1
    coef = np.zeros((nb, 3))
    icpt = np.zeros(nb)
    used = [nm for nm in feature_names if nm in names]
    for b in range(nb):
        c = np.asarray(reg.LRs[b].coef_, dtype=np.float64).ravel()
        for nm, v in zip(used, c):
            if nm not in FEATURE_NAMES[:3]:
                raise NotImplementedError("regression feature %r is not available on the device" % nm)
            coef[b, FEATURE_NAMES.index(nm)] = v
        icpt[b] = float(reg.LRs[b].intercept_)
    return bins, coef, icpt


# ---------------------------------------------------------------------------------------------
# error predictor (annchor/error_predictors.py:18-67)
# ---------------------------------------------------------------------------------------------
class SimpleStratifiedErrorRegression:
    def __init__(self, partition_feature_name="double anchor distance", n_partitions=7):
        self.n_partitions = n_partitions
        self.partition_feature_name = partition_feature_name
        self.labels = range(n_partitions)

    def fit(self, sample_features, feature_names, sample_error, sample_bins=None):
        f = sample_features[:, feature_names.index(self.partition_feature_name)]
        if sample_bins is None:
            n = f.shape[0]
            q1 = np.partition(f, int(n / 100))[int(n / 100)]
            q3 = np.partition(f, int(99 * n / 100))[int(99 * n / 100)]
            self.partition_bins = np.hstack([-np.inf, np.linspace(q1, q3, self.n_partitions - 1), np.inf])
        else:
            self.n_partitions = sample_bins.shape[0] - 1
            self.partition_bins = sample_bins
        self.labels = range(self.n_partitions)
        self.errs = {}
        for b in range(self.n_partitions):
            m = (f >= self.partition_bins[b]) & (f <= self.partition_bins[b + 1])
            self.errs[b] = np.sort(sample_error[m])

    def predict(self, features, feature_names):
        f = features[:, feature_names.index(self.partition_feature_name)]
        labels = np.full(features.shape[0], -1, dtype=np.int64)
        for b in range(self.n_partitions):
            labels[(f >= self.partition_bins[b]) & (f <= self.partition_bins[b + 1])] = b
        return labels


def error_device_spec(ep, n_bins=None):
    """(errs_flat, eptr) of a fitted error predictor (ours or annchor/error_predictors.py:47-54)."""
    if not hasattr(ep, "errs"):
        raise NotImplementedError("error predictors must expose per-bin sorted error tables `.errs`")
    nb = len(ep.errs)
    if n_bins is not None and nb != n_bins:
        raise ValueError("the error predictor has %d bins, the regression %d: the device scores with one "
                         "partition for both (annchor/annchor.py:150-161 builds both from the sampler's bins)"
                         % (nb, n_bins))
    tabs = [np.asarray(ep.errs[b], dtype=np.float64) for b in range(nb)]
    eptr = np.zeros(nb + 1, dtype=np.int64)
    np.cumsum([len(t) for t in tabs], out=eptr[1:])
    if any(len(t) == 0 for t in tabs):
        raise Exception("an error-predictor bin is empty (no samples fell into it)")
    return np.concatenate(tabs), eptr
