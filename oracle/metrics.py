"""Oracle metrics: pair-batch evaluators with the reference's get_exact_ijs
contract  ``out[p] = f(X[i_p], X[j_p])``  (annchor/utils.py:110-177).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).
"""
import numpy as np

from .clib import lib, ptr


def pack_strings(X):
    """Pack a sequence of str into (uint8 chars, int64 offsets[n+1]).

    The reference's strings are numpy '<U594' over a-z (annchor/datasets.py:91-125);
    code points must fit one byte here (the DP itself is alphabet-agnostic).
    """
    enc = [s.encode("latin-1") for s in X]
    offs = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in enc], out=offs[1:])
    chars = np.frombuffer(b"".join(enc), dtype=np.uint8).copy()
    return chars, offs


def levenshtein(a, b):
    """Unit-cost edit distance of two str (annchor/distances.py:16-20)."""
    ea, eb = a.encode("latin-1"), b.encode("latin-1")
    ba = np.frombuffer(ea, dtype=np.uint8) if ea else np.zeros(0, np.uint8)
    bb = np.frombuffer(eb, dtype=np.uint8) if eb else np.zeros(0, np.uint8)
    return int(lib().orc_lev(ptr(ba), len(ea), ptr(bb), len(eb)))


def _ij(IJ):
    IJ = np.ascontiguousarray(IJ, dtype=np.int64).reshape(-1, 2)
    return IJ, np.empty(IJ.shape[0], dtype=np.float64)


class PairMetric:
    """Callable IJ -> float64 distances for one dataset."""

    def __init__(self, X, name, **kw):
        self.name = name
        self.n = len(X)
        L = lib()
        if name == "levenshtein":
            self.chars, self.offs = pack_strings(X)
            self._call = lambda ij, n, out: L.orc_lev_pairs(
                ptr(self.chars), ptr(self.offs), ptr(ij), n, ptr(out))
        elif name in ("euclidean", "cosine"):
            X = np.asarray(X)
            if X.dtype == np.float32:
                self.X = np.ascontiguousarray(X)
                fn = getattr(L, "orc_%s_pairs_f32" % ("euclid" if name == "euclidean" else "cosine"))
            else:
                self.X = np.ascontiguousarray(X, dtype=np.float64)
                fn = getattr(L, "orc_%s_pairs_f64" % ("euclid" if name == "euclidean" else "cosine"))
            d = self.X.shape[1]
            self._call = lambda ij, n, out: fn(ptr(self.X), d, ptr(ij), n, ptr(out))
        elif name == "wasserstein1d":
            self.X = np.ascontiguousarray(X, dtype=np.float64)
            nb = self.X.shape[1]
            self._call = lambda ij, n, out: L.orc_w1_pairs_f64(ptr(self.X), nb, ptr(ij), n, ptr(out))
        elif name == "wasserstein":
            # general ground cost (annchor/utils.py:75-86, kantorovich(x, y, cost=M)): exact OT
            self.X = np.ascontiguousarray(X, dtype=np.float64)
            nb = self.X.shape[1]
            self.M = np.ascontiguousarray(kw["cost_matrix"], dtype=np.float64)
            assert self.M.shape == (nb, nb)

            def call(ij, n, out):
                rc = L.orc_ot_pairs_f64(ptr(self.X), nb, ptr(self.M), ptr(ij), n, ptr(out))
                assert rc == 0, "too many bins for the oracle's OT solver"
            self._call = call
        else:
            raise ValueError("unknown oracle metric %r" % (name,))

    def __call__(self, IJ):
        ij, out = _ij(IJ)
        if ij.shape[0]:
            self._call(ij, ij.shape[0], out)
        return out


def wasserstein1d_numpy(x, y):
    """1-D Wasserstein between two histograms on the unit-spaced bin index:
    sum |CDF_x - CDF_y| on unit-mass normalisations (what
    kantorovich(x, y, cost=|a-b|) of annchor/utils.py:82-84 evaluates to)."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    return float(np.abs(np.cumsum(x / x.sum()) - np.cumsum(y / y.sum())).sum())


def kantorovich_lp(x, y, M):
    """kantorovich(x, y, cost=M) (annchor/utils.py:82-84) as a linear programme (scipy HiGHS): an exact solver
    independent of the augmenting-path one in oracle.c, used to pin it on a few pairs."""
    from scipy.optimize import linprog
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    si, tj = np.nonzero(x > 0)[0], np.nonzero(y > 0)[0]
    a, b = x[si] / x.sum(), y[tj] / y.sum()
    m, n = len(si), len(tj)
    c = np.asarray(M, dtype=np.float64)[np.ix_(si, tj)].ravel()
    A = np.zeros((m + n, m * n))
    for k in range(m):
        A[k, k * n:(k + 1) * n] = 1.0
    for l in range(n):
        A[m + l, l::n] = 1.0
    r = linprog(c, A_eq=A[:-1], b_eq=np.concatenate([a, b])[:-1], bounds=(0, None), method="highs")
    assert r.status == 0, r.message
    return float(r.fun)
