"""Device-arithmetic mode of the oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``OracleAnnchor`` (oracle/pipeline.py) restates the reference's ``fit()`` in the reference's own
arithmetic: float64 everywhere, ties left to numpy's unstable ``argpartition`` / ``argsort``.
The CUDA product documents two deliberate deviations:

  * the Theta(N^2) stages (bounds, dad, prediction, clip, thresholds, probabilities) run in
    float32 from a float32 copy of D, the prediction as one fused multiply-add chain;
  * every tie the reference leaves to an unstable sort is resolved by a fixed rule: selection
    ties at the cut by the smallest ``tie_key(pair, salt)``, top-k ties by neighbour id,
    guarantee_nmin lists by (value, id).

``OracleAnnchorF32`` is the same pipeline with exactly those two rules applied (same stage
order, same reference semantics otherwise: annchor/annchor.py:345-530, utils.py:274-429,581-621),
so that every stage of the device fit -- thresholds, forced pairs, selected and look-ahead sets,
tightened bounds, sample, model, final graph -- can be compared for EQUALITY instead of overlap
percentages.  Metric values are an input here (``pair_fn``; the tests pass the device's own
metric kernel, whose parity with the reference metrics is pinned separately in
tests/test_metrics_gpu.py), and so may be the anchors (A, D).
"""
import numpy as np

from .clib import lib, ptr
from .pipeline import (OracleAnnchor, NothingToSample, fit_stratified_linear,
                       predict_stratified_linear, row_kth)

M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _u64(x):
    return np.asarray(x).astype(np.uint64)


def mix64(x):
    """splitmix64 finaliser (annchor_b200/csrc/index.cuh: mix64)."""
    x = _u64(x).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
    return x


def hash_pair32(i, j, seed):
    """index.cuh: hash_pair32"""
    i = np.asarray(i).astype(np.uint32)
    j = np.asarray(j).astype(np.uint32)
    with np.errstate(over="ignore"):
        h = (i * np.uint32(0x9E3779B1)) ^ (j * np.uint32(0x85EBCA77)) ^ np.uint32(seed & 0xFFFFFFFF)
        h ^= h >> np.uint32(15)
        h *= np.uint32(0x2C1B3C6D)
        h ^= h >> np.uint32(12)
        h *= np.uint32(0x297A2D39)
        h ^= h >> np.uint32(15)
    return h


def tie_key(lo, hi, salt):
    """index.cuh: tie_key -- smaller wins at a selection cut."""
    salt = int(salt) & 0xFFFFFFFFFFFFFFFF
    key = (_u64(lo) << np.uint64(32)) | _u64(hi)
    low = mix64(key ^ np.uint64(salt)) & np.uint64(0xFFFFFFFF)
    return (hash_pair32(lo, hi, salt & 0xFFFFFFFF).astype(np.uint64) << np.uint64(32)) | low


def select_salt(n_selects, salt0=0):
    """annb_index_select: the salt changes with every selection of an index."""
    return int(mix64(np.uint64((0x9E3779B97F4A7C15 * n_selects + salt0) & 0xFFFFFFFFFFFFFFFF)))


class OracleAnnchorF32(OracleAnnchor):
    def __init__(self, X, metric, A=None, D=None, tie_salt0=0, **kw):
        super().__init__(X, metric, **kw)
        if not self.is_metric:
            raise NotImplementedError("device-arithmetic mode covers is_metric=True")
        self._A_in = None if A is None else np.asarray(A, dtype=np.int64)
        self._D_in = None if D is None else np.ascontiguousarray(D, dtype=np.float64)
        self.tie_salt0 = tie_salt0
        self.n_selects = 0
        self.n_forced = 0

    # -- anchors: optionally injected (device values) --------------------------------------
    def get_anchors(self):
        if self._D_in is None:
            return super().get_anchors()
        self.A = self._A_in
        self.D = self._D_in
        self.evals += self.n_anchors * self.nx

    # -- features: float64 for the host side (sampler / regression fit), float32 for the sweeps
    def get_features(self):
        super().get_features()
        P = self.IJs.shape[0]
        self.D32 = np.ascontiguousarray(self.D.astype(np.float32))
        self.cA = np.ascontiguousarray(np.argmin(self.D, axis=1).astype(np.int32))
        self.lb32 = np.empty(P, np.float32)
        self.ub32 = np.empty(P, np.float32)
        self.s2 = np.empty(P, np.float32)
        lib().orc_f32_features(ptr(self.IJs), P, ptr(self.D32), self.n_anchors, ptr(self.cA),
                               ptr(self.lb32), ptr(self.ub32), ptr(self.s2))
        self.known = np.zeros(P, dtype=bool)          # KIND_KNOWN entries of the device store
        self.known32 = np.zeros(P, dtype=np.float32)
        self.tight = np.zeros(P, dtype=bool)          # KIND_TIGHT entries
        self.anchor_pair = ~self.not_computed_mask.copy()
        # value an anchor pair contributes to the final graph: D64[other, slot of the anchor]
        slot = np.full(self.nx, -1, dtype=np.int64)
        for k, a in enumerate(self.A):
            slot[a] = k
        self.slot = slot

    def get_sample(self):
        super().get_sample()
        self.known[self.sample_ixs] = True
        self.known32[self.sample_ixs] = self.sample_y.astype(np.float32)

    # -- model -----------------------------------------------------------------------------
    def _device_model(self):
        nb = self.sample_bins.shape[0] - 1
        edge = self.sample_bins.astype(np.float32)
        e2 = np.full(8, np.inf, dtype=np.float32)
        for b in range(1, nb):
            e2[b] = np.float32(2.0) * edge[b]
        cf = np.zeros((8, 4), dtype=np.float32)
        for b in range(nb):
            c = self.coef[b].astype(np.float32)
            cf[b] = (c[0], c[1], np.float32(0.5) * c[2], np.float32(self.icpt[b]))
        return e2, cf

    def fit_predict_regression(self):
        F = self.sample_features[:, 2]
        self.coef, self.icpt = fit_stratified_linear(F, self.sample_features[:, :3], self.sample_y,
                                                     self.sample_bins)
        self.sample_predict = predict_stratified_linear(F, self.sample_features[:, :3], self.sample_bins,
                                                        self.coef, self.icpt)
        P = self.IJs.shape[0]
        e2, cf = self._device_model()
        self.pred32 = np.empty(P, np.float32)
        self.bin8 = np.empty(P, np.int8)
        self.label8 = np.empty(P, np.int8)
        lib().orc_f32_predict(ptr(self.lb32), ptr(self.ub32), ptr(self.s2), P, ptr(e2), ptr(cf),
                              ptr(self.pred32), ptr(self.bin8), ptr(self.label8))
        # RefineApprox as annchor.py:374-380 leaves it: exact where computed, prediction elsewhere
        # (anchor pairs: lb == ub == D so the clipped prediction IS the distance)
        self.RA32 = np.where(self.known, self.known32, self.pred32).astype(np.float32)
        self.RefineApprox = self.RA32.astype(np.float64)
        self.pred = self.pred32

    def fit_predict_errors(self):
        F = self.sample_features[:, 2]
        err = self.sample_y - self.sample_predict
        bins = self.sample_bins
        self.errs = [np.sort(err[(F >= bins[b]) & (F <= bins[b + 1])]) for b in range(bins.shape[0] - 1)]
        self.errs32 = [e.astype(np.float32) for e in self.errs]
        self.errors = self.label8.astype(np.int64)

    # -- select / refine -------------------------------------------------------------------
    def select_refine_candidate_pairs(self, w, it):
        nn = self.n_neighbors
        self.thresh = row_kth(self.RefineApprox, self.row_ptr, self.row_pairs, nn)  # float32 values
        if it == 0:
            self.guarantee_nmin(3 * nn // 2)  # reference algorithm on the float32 values
        ncm = self.not_computed_mask
        forced = ncm & (self.RefineApprox == -1)
        if it == 0:
            self.n_forced = int(forced.sum())
            self.forced_pairs = self.IJs[forced]
        v32 = self.RefineApprox.astype(np.float32)
        th32 = self.thresh.astype(np.float32)
        I, J = self.IJs[:, 0], self.IJs[:, 1]
        p32 = (np.maximum(th32[I], th32[J]) - v32)[ncm]
        lab = self.errors[ncm]
        prob = np.zeros(p32.shape[0])
        for b, e in enumerate(self.errs32):
            m = lab == b
            prob[m] = np.searchsorted(e, p32[m], side="left") / float(len(e))
        self.prob = prob
        n_refine = max(int((self.p_work * self.N - self.na - self.n_samples) * w) + 1, 0)
        self.n_refine = n_refine
        n_nc = prob.shape[0]
        sel_target = min(n_refine, n_nc)
        tot_target = n_nc if n_refine >= n_nc else min(n_refine * self.lookahead, n_nc)
        self.n_selects += 1
        salt = select_salt(self.n_selects, self.tie_salt0)
        back = np.arange(ncm.shape[0])[ncm]
        tk = tie_key(I[back], J[back], salt)
        order = np.lexsort((tk, -prob))  # probability descending, then the smaller tie key
        mapback = back[order[:sel_target]]
        self.nextback = back[order[sel_target:tot_target]]
        self.mapback = mapback
        exact = self.pair_fn(self.IJs[mapback])
        self.evals += exact.shape[0]
        self.known[mapback] = True
        self.known32[mapback] = exact.astype(np.float32)
        self.RA32[mapback] = self.known32[mapback]
        self.RefineApprox[mapback] = self.known32[mapback]
        ncm[mapback] = False

    # -- tightening (annchor.py:475-512 / utils.py:304-352 in float32 over the KNOWN entries) -----
    def known_lists32(self):
        k = np.nonzero(self.known)[0]
        ij = self.IJs[k]
        d = self.known32[k]
        src = np.concatenate([ij[:, 0], ij[:, 1]])
        dst = np.concatenate([ij[:, 1], ij[:, 0]])
        dd = np.concatenate([d, d])
        o = np.lexsort((dst, src))
        kptr = np.zeros(self.nx + 1, dtype=np.int64)
        np.cumsum(np.bincount(src, minlength=self.nx), out=kptr[1:])
        return kptr, np.ascontiguousarray(dst[o]), np.ascontiguousarray(dd[o])

    def update_anchor_points(self):
        mb = self.nextback
        self.n_tightened = 0
        if mb.shape[0] == 0:
            return
        kptr, kids, kds = self.known_lists32()
        ij = np.ascontiguousarray(self.IJs[mb])
        lb = np.empty(mb.shape[0], np.float32)
        ub = np.empty(mb.shape[0], np.float32)
        lib().orc_f32_update_bounds(ptr(ij), mb.shape[0], ptr(kptr), ptr(kids), ptr(kds), ptr(lb), ptr(ub))
        l0, u0 = self.lb32[mb], self.ub32[mb]
        improved = (lb > l0) | (ub < u0)
        self.lb32[mb] = np.maximum(lb, l0)
        self.ub32[mb] = np.minimum(ub, u0)
        imp = mb[improved]
        self.tight[imp] = True
        self.n_tightened = int(improved.sum())
        self.tightened_pairs = self.IJs[imp]
        # what the exact-mode sampler of the product sees: float64 anchor bounds overlaid with the
        # (float32) tightened ones
        self.features[imp, 0] = np.maximum(self.features[imp, 0], self.lb32[imp].astype(np.float64))
        self.features[imp, 1] = np.minimum(self.features[imp, 1], self.ub32[imp].astype(np.float64))

    # -- final graph: computed pairs only, ordered by (distance, neighbour id) ------------------
    def get_ann(self):
        nx, nn = self.nx, self.n_neighbors
        I, J = self.IJs[:, 0], self.IJs[:, 1]
        comp = self.known | self.anchor_pair
        idx = np.full((nx, nn), -1, dtype=np.int64)
        dist = np.full((nx, nn), np.inf)
        idx[:, 0] = np.arange(nx)
        dist[:, 0] = 0.0
        for i in range(nx):
            pr = self.row(i)
            pr = pr[comp[pr]]
            other = np.where(I[pr] == i, J[pr], I[pr])
            d = self.known32[pr].astype(np.float64)
            ap = ~self.known[pr]
            if ap.any():
                # anchor pairs take the float64 anchor distance: this row's own anchor slot if the
                # row is an anchor, else the other endpoint's
                if self.slot[i] >= 0:
                    d[ap] = self.D[other[ap], self.slot[i]]
                else:
                    d[ap] = self.D[i, self.slot[other[ap]]]
            o = np.lexsort((other, d))[:nn - 1]
            m = o.shape[0]
            idx[i, 1:1 + m] = other[o]
            dist[i, 1:1 + m] = d[o]
            if m < nn - 1:
                # utils.py:415-428: not-computed candidates rank behind the computed ones by their
                # RefineApprox (a prediction), which is also what is emitted
                pr = self.row(i)
                pr = pr[~comp[pr]]
                oth = np.where(I[pr] == i, J[pr], I[pr])
                v = self.RA32[pr].astype(np.float64)
                o2 = np.lexsort((oth, v))[:nn - 1 - m]
                idx[i, 1 + m:1 + m + o2.shape[0]] = oth[o2]
                dist[i, 1 + m:1 + m + o2.shape[0]] = v[o2]
        self.neighbor_graph = (idx, dist)

    # -- query (annchor/annchor.py:643-683, annchor/query_functions.py:10-212) in the same arithmetic ----
    def query(self, q_pair_fn, nq, nn=15, p_work=0.3):
        """q_pair_fn(IJ) -> metric(X[i], Q[j]) for (i, j) rows of IJ.  Returns (ngi, ngd) of shape (nq, nn)."""
        nx, na = self.nx, self.n_anchors
        limit = ((nq * nn * 3) // 2 - 1 + na * nq) / (nq * nx)
        p_work = max(p_work, limit)
        # query -> anchor distances (query_functions.py:10-15)
        IJa = np.array([[self.A[a], j] for j in range(nq) for a in range(na)], dtype=np.int64)
        QD = q_pair_fn(IJa).reshape(nq, na)
        evals = nq * na
        # candidates: queries and points sharing >= loc_thresh of their `locality` nearest anchors (:18-38)
        sidq = np.argsort(QD, axis=1, kind="stable")[:, :self.locality]
        Mq = np.zeros((nq, na), dtype=np.int32)
        np.put_along_axis(Mq, sidq, 1, axis=1)
        Mx = np.zeros((nx, na), dtype=np.int32)
        np.put_along_axis(Mx, self.sid, 1, axis=1)
        cand = (Mq @ Mx.T) >= self.loc_thresh                      # (nq, nx)
        jj, ii = np.nonzero(cand)
        P = ii.shape[0]
        # features in float32 on the data set "X followed by Q" (query_functions.py:70-130)
        Dc = np.ascontiguousarray(np.vstack([self.D, QD]).astype(np.float32))
        cAc = np.ascontiguousarray(np.concatenate([np.argmin(self.D, axis=1), np.argmin(QD, axis=1)]).astype(np.int32))
        ij = np.ascontiguousarray(np.stack([ii, nx + jj], axis=1).astype(np.int64))
        lb, ub, s2 = np.empty(P, np.float32), np.empty(P, np.float32), np.empty(P, np.float32)
        lib().orc_f32_features(ptr(ij), P, ptr(Dc), na, ptr(cAc), ptr(lb), ptr(ub), ptr(s2))
        e2, cf = self._device_model()
        RA = np.empty(P, np.float32)
        lab = np.empty(P, np.int8)
        lib().orc_f32_predict(ptr(lb), ptr(ub), ptr(s2), P, ptr(e2), ptr(cf), ptr(RA), None, ptr(lab))
        computed = self.slot[ii] >= 0                              # np.isin(IJs[:, 0], ann.A)
        # rows of the rectangle (jj is sorted: np.nonzero is row-major)
        rptr = np.zeros(nq + 1, dtype=np.int64)
        np.cumsum(np.bincount(jj, minlength=nq), out=rptr[1:])
        thresh = np.array([np.sort(RA[rptr[j]:rptr[j + 1]])[nn] for j in range(nq)], dtype=np.float32)
        nmin = 3 * nn // 2
        for j in range(nq):                                        # guarantee_nmin, row by row (rows are disjoint)
            sl = slice(rptr[j], rptr[j + 1])
            m = ~computed[sl]
            todo = nmin - int((~m).sum())
            if todo > 0:
                vals = RA[sl][m]
                kth = np.partition(vals, todo)[todo]
                idx = np.nonzero(m)[0][vals < kth] + rptr[j]
                RA[idx] = -1.0
        ncm = ~computed
        p32 = (thresh[jj] - RA)[ncm]
        labn = lab[ncm].astype(np.int64)
        prob = np.zeros(p32.shape[0])
        for b, e in enumerate(self.errs32):
            m = labn == b
            prob[m] = np.searchsorted(e, p32[m], side="left") / float(len(e))
        n_refine = int(p_work * nq * nx - na * nq) + 1
        n_refine = max(0, min(n_refine, prob.shape[0]))
        self.n_selects += 1
        salt = select_salt(self.n_selects, self.tie_salt0)
        back = np.nonzero(ncm)[0]
        tk = tie_key(ii[back], nx + jj[back], salt)
        sel = back[np.lexsort((tk, -prob))[:n_refine]]
        exact = q_pair_fn(np.stack([ii[sel], jj[sel]], axis=1)).astype(np.float32)
        evals += sel.shape[0]
        RA[sel] = exact
        computed[sel] = True
        ngi = np.full((nq, nn), -1, dtype=np.int64)
        ngd = np.full((nq, nn), np.inf)
        for j in range(nq):
            sl = slice(rptr[j], rptr[j + 1])
            pts, v, c = ii[sl], RA[sl].astype(np.float64), computed[sl]
            o = np.lexsort((pts[c], v[c]))[:nn]
            m = o.shape[0]
            ngi[j, :m] = pts[c][o]
            ngd[j, :m] = v[c][o]
            if m < nn:
                o2 = np.lexsort((pts[~c], v[~c]))[:nn - m]
                ngi[j, m:m + o2.shape[0]] = pts[~c][o2]
                ngd[j, m:m + o2.shape[0]] = v[~c][o2]
        self.query_evals = evals
        return ngi, ngd

    def fit(self):
        self.get_anchors()
        self._t("A", self.A)
        self.get_locality()
        self.get_features()
        for it in range(self.niters):
            try:
                self.get_sample()
            except NothingToSample as err:
                if it == 0:
                    raise ValueError("Sampler raised NothingToSample on first iteration.") from err
                break
            self._t("sample_ijs%d" % it, self.IJs[self.sample_ixs])
            self._t("sample_bins%d" % it, self.sample_bins)
            self._t("sample_features%d" % it, self.sample_features)
            self.fit_predict_regression()
            self._t("coef%d" % it, self.coef)
            self.fit_predict_errors()
            self.select_refine_candidate_pairs(1 / self.niters, it)
            self._t("thresh%d" % it, self.thresh)
            self._t("selected%d" % it, self.IJs[self.mapback])
            self._t("next%d" % it, self.IJs[self.nextback])
            if it == 0:
                self._t("n_forced", self.n_forced)
            if it < self.niters - 1:
                self.update_anchor_points()
                self._t("n_tightened%d" % it, self.n_tightened)
        self.get_ann()
        return self
