"""Stage operators with the reference's names and array conventions, running on the B200.

Each function is a drop-in for the numba function (or predict() method) it cites: same
arguments where they are arrays, same outputs, float64 / int64 host arrays in and out.  The
per-point pair index ``I`` (a numba typed Dict in the reference) is passed as a CSR pair
``(row_ptr, row_pairs)``; ``csr_from_I`` converts a dict-like."""
import numpy as np

from . import _lib
from ._lib import check, ptr, as_c
from .core import default_context


def _ctx(ctx):
    return ctx or default_context()


def csr_from_I(I, nx):
    """numba Dict / dict {i: int64[:]} -> (row_ptr, row_pairs)."""
    lens = np.array([len(I[i]) for i in range(nx)], dtype=np.int64)
    row_ptr = np.zeros(nx + 1, dtype=np.int64)
    np.cumsum(lens, out=row_ptr[1:])
    row_pairs = np.concatenate([np.asarray(I[i], dtype=np.int64) for i in range(nx)]) \
        if nx else np.zeros(0, np.int64)
    return row_ptr, row_pairs


def get_bounds_njit_ijs(IJs, D, ctx=None):
    """annchor/utils.py:274-301 -> float64 (n, 2) [lower, upper]."""
    IJs = as_c(IJs, np.int64).reshape(-1, 2)
    D = as_c(D, np.float64)
    out = np.empty((IJs.shape[0], 2), dtype=np.float64)
    check(_lib.load().annb_bounds_ijs(_ctx(ctx).handle, ptr(IJs), IJs.shape[0], ptr(D), D.shape[0],
                                      D.shape[1], ptr(out)))
    return out


def get_dad_ijs(IJs, D, ctx=None):
    """annchor/utils.py:355-380 -> float64 (n,)."""
    IJs = as_c(IJs, np.int64).reshape(-1, 2)
    D = as_c(D, np.float64)
    out = np.empty(IJs.shape[0], dtype=np.float64)
    check(_lib.load().annb_dad_ijs(_ctx(ctx).handle, ptr(IJs), IJs.shape[0], ptr(D), D.shape[0],
                                   D.shape[1], ptr(out)))
    return out


def update_bounds(IJs, kptr, kids, kds, ctx=None):
    """annchor/utils.py:326-352 with the per-point known lists (dis, ds) as CSR."""
    IJs = as_c(IJs, np.int64).reshape(-1, 2)
    kptr, kids, kds = as_c(kptr, np.int64), as_c(kids, np.int64), as_c(kds, np.float64)
    out = np.empty((IJs.shape[0], 2), dtype=np.float64)
    check(_lib.load().annb_update_bounds(_ctx(ctx).handle, ptr(IJs), IJs.shape[0], ptr(kptr),
                                         ptr(kids), ptr(kds), kptr.shape[0] - 1, ptr(out)))
    return out


def predict_stratified(features, bins, coef, icpt, ctx=None):
    """SimpleStratifiedLinearRegression.predict (annchor/regressors.py:71-103) and the clip of
    annchor/annchor.py:359-363 -> (pred_raw, pred_clipped)."""
    features = as_c(features, np.float64)
    bins, coef, icpt = as_c(bins, np.float64), as_c(coef, np.float64), as_c(icpt, np.float64)
    n = features.shape[0]
    raw, clp = np.empty(n), np.empty(n)
    check(_lib.load().annb_predict_stratified(_ctx(ctx).handle, ptr(features), n, ptr(bins),
                                              ptr(coef), ptr(icpt), bins.shape[0] - 1, ptr(raw),
                                              ptr(clp)))
    return raw, clp


def error_labels(feature, bins, ctx=None):
    """SimpleStratifiedErrorRegression.predict (annchor/error_predictors.py:56-67)."""
    feature, bins = as_c(feature, np.float64), as_c(bins, np.float64)
    out = np.empty(feature.shape[0], dtype=np.int64)
    check(_lib.load().annb_error_labels(_ctx(ctx).handle, ptr(feature), feature.shape[0], ptr(bins),
                                        bins.shape[0] - 1, ptr(out)))
    return out


def get_probs(p, labels, errs, ctx=None):
    """annchor/utils.py:581-589; errs = sequence of sorted float64 arrays indexed by label."""
    p, labels = as_c(p, np.float64), as_c(labels, np.int64)
    eptr = np.zeros(len(errs) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in errs], out=eptr[1:])
    flat = as_c(np.concatenate([np.asarray(e, dtype=np.float64) for e in errs]), np.float64)
    out = np.empty(p.shape[0], dtype=np.float64)
    check(_lib.load().annb_probs(_ctx(ctx).handle, ptr(p), ptr(labels), p.shape[0], ptr(flat),
                                 ptr(eptr), len(errs), ptr(out)))
    return out


def row_kth(RA, row_ptr, row_pairs, k, ctx=None):
    """thresh loop of annchor/annchor.py:399-404."""
    RA, row_ptr, row_pairs = as_c(RA, np.float64), as_c(row_ptr, np.int64), as_c(row_pairs, np.int64)
    nx = row_ptr.shape[0] - 1
    out = np.empty(nx, dtype=np.float64)
    check(_lib.load().annb_row_kth(_ctx(ctx).handle, ptr(RA), RA.shape[0], ptr(row_ptr),
                                   ptr(row_pairs), nx, int(k), ptr(out)))
    return out


def get_nn(nx, nn, RA, IJs, row_ptr, row_pairs, not_computed_mask, ctx=None):
    """annchor/utils.py:383-429 -> (ngi int64 (nx, nn-1), ngd float64 (nx, nn-1))."""
    RA, IJs = as_c(RA, np.float64), as_c(IJs, np.int64)
    row_ptr, row_pairs = as_c(row_ptr, np.int64), as_c(row_pairs, np.int64)
    m8 = as_c(not_computed_mask, np.uint8)
    ngi = np.empty((nx, nn - 1), dtype=np.int64)
    ngd = np.empty((nx, nn - 1), dtype=np.float64)
    check(_lib.load().annb_get_nn(_ctx(ctx).handle, nx, nn, ptr(RA), ptr(IJs), RA.shape[0],
                                  ptr(row_ptr), ptr(row_pairs), ptr(m8), ptr(ngi), ptr(ngd)))
    return ngi, ngd
