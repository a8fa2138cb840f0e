"""Host-side pieces of the path that need no GPU: the reference's tie-aware graph metric
(annchor/annchor.py:1026-1066, reference test annchor/tests/test_annchor.py:15-32), the plug-in
objects (annchor/regressors.py, error_predictors.py, samplers.py) against the oracle's restatement,
and the device specs the streaming index consumes."""
import numpy as np
import pytest

from annchor_b200.annchor import compare_neighbor_graphs
from annchor_b200 import plugins
from oracle import pipeline as P


def _graph(n=50, k=10, seed=0):
    rng = np.random.default_rng(seed)
    d = np.sort(rng.random((n, k)), axis=1)
    d[:, 0] = 0
    idx = np.stack([rng.permutation(n)[:k] for _ in range(n)])
    idx[:, 0] = np.arange(n)
    return idx, d


def test_compare_neighbor_graphs_reference_cases():
    g = _graph()
    assert compare_neighbor_graphs(g, g, 10) == 0
    # injected errors are counted exactly (reference test: self = 0, k wrong entries = k)
    idx, d = g[0].copy(), g[1].copy()
    d[3, 5] += 1.0
    d[7, 2] += 1.0
    d[7, 9] += 1.0
    assert compare_neighbor_graphs(g, (idx, d), 10) == 3
    # tie-aware: permuting equal distances is not an error, and indices do not matter
    d2 = g[1].copy()
    d2[4, 3] = d2[4, 4]
    a = (g[0], d2)
    b_idx = g[0].copy()
    b_idx[4, [3, 4]] = b_idx[4, [4, 3]]
    assert compare_neighbor_graphs(a, (b_idx, d2), 10) == 0
    # rounding to 3 decimals (annchor.py:1058-1061)
    d3 = np.round(g[1], 2)          # values on a 0.01 grid: +1e-5 never crosses a rounding boundary
    d3b = d3.copy()
    d3b[:, 1:] += 1e-5
    assert compare_neighbor_graphs((g[0], d3), (g[0], d3b), 10) == 0
    d3b[:, 1:] += 1e-3
    assert compare_neighbor_graphs((g[0], d3), (g[0], d3b), 10) > 0
    # only the first n_neighbors columns are compared
    d4 = g[1].copy()
    d4[:, 9] += 5
    assert compare_neighbor_graphs(g, (g[0], d4), 9) == 0
    assert compare_neighbor_graphs(g, (g[0], d4), 10) == 50
    # same as the oracle's restatement
    assert compare_neighbor_graphs(g, (idx, d), 10) == P.compare_neighbor_graphs(g, (idx, d), 10)


def _sample(n=4000, seed=1):
    rng = np.random.default_rng(seed)
    lb = rng.random(n) * 5
    ub = lb + rng.random(n) * 5
    dad = (lb + ub) / 2 + rng.normal(size=n) * 0.3
    y = 0.3 * lb + 0.5 * ub + 0.2 * dad + rng.normal(size=n) * 0.1
    feats = np.stack([lb, ub, dad, np.zeros(n)], axis=1)
    return feats, y


def test_regression_and_error_predictor_match_oracle():
    feats, y = _sample()
    names = list(plugins.FEATURE_NAMES)
    bins, _ = P.stratified_partition(feats[:, 2], 5000)
    reg = plugins.SimpleStratifiedLinearRegression()
    reg.fit(feats, names, y, sample_bins=bins)
    coef, icpt = P.fit_stratified_linear(feats[:, 2], feats[:, :3], y, bins)
    b2, c2, i2 = plugins.regression_device_spec(reg, names)
    np.testing.assert_array_equal(b2, bins)
    np.testing.assert_allclose(c2, coef, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(i2, icpt, rtol=1e-12, atol=1e-12)
    pred = reg.predict(feats, names)
    np.testing.assert_allclose(pred, P.predict_stratified_linear(feats[:, 2], feats[:, :3], bins, coef, icpt),
                               rtol=1e-12, atol=1e-12)
    # bin edge convention (lo, hi]: a value exactly on an edge belongs to the lower bin
    edge = feats[:1].copy()
    edge[0, 2] = bins[3]
    assert np.allclose(reg.predict(edge, names), edge[0, :3] @ coef[2] + icpt[2])
    # error predictor: closed intervals, later bins win; sorted residual tables
    ep = plugins.SimpleStratifiedErrorRegression()
    ep.fit(feats, names, y - pred, sample_bins=bins)
    np.testing.assert_array_equal(ep.predict(feats, names), P.error_labels(feats[:, 2], bins))
    assert ep.predict(edge, names)[0] == 3
    errs, eptr = plugins.error_device_spec(ep)
    assert eptr[0] == 0 and eptr[-1] == errs.shape[0] and len(eptr) == len(bins)
    for b in range(len(bins) - 1):
        tab = errs[eptr[b]:eptr[b + 1]]
        assert np.all(np.diff(tab) >= 0)
        m = (feats[:, 2] >= bins[b]) & (feats[:, 2] <= bins[b + 1])
        np.testing.assert_array_equal(tab, np.sort((y - pred)[m]))


def test_sampler_partition_matches_oracle_and_falls_back():
    rng = np.random.default_rng(3)
    f = rng.random(100000)
    s = plugins.SimpleStratifiedSampler()
    bins, ns = s.get_partition(f, 5000)
    ob, ons = P.stratified_partition(f, 5000)
    np.testing.assert_array_equal(bins, ob)
    assert ns == ons == 5000
    assert bins[0] == -np.inf and bins[-1] == np.inf and len(bins) == 8
    # too few points for 1 % tails: 10 % tails, then a reduced sample (annchor/samplers.py:119-133)
    f2 = rng.random(3000)
    bins2, ns2 = s.get_partition(f2, 5000)
    ob2, ons2 = P.stratified_partition(f2, 5000)
    np.testing.assert_array_equal(bins2, ob2)
    assert ns2 == ons2 == 300 * 7


def test_device_specs_reject_opaque_plugins():
    class Opaque:
        def fit(self, *a, **k):
            pass

        def predict(self, f, names):
            return np.zeros(len(f))

    with pytest.raises(NotImplementedError):
        plugins.regression_device_spec(Opaque(), list(plugins.FEATURE_NAMES))
    with pytest.raises(NotImplementedError):
        plugins.error_device_spec(Opaque())


def test_sampler_bins_are_the_half_open_intervals_of_the_reference():
    """bin_of == np.digitize == the reference's (lo <= x) & (x < hi) masks (annchor/utils.py:547-549), values on the
    edges and +-inf edges included."""
    from annchor_b200.plugins import bin_of
    rng = np.random.default_rng(0)
    bins = np.hstack([-np.inf, np.linspace(0.2, 0.8, 6), np.inf])
    v = np.concatenate([rng.random(5000), bins[1:-1], bins[1:-1] - 1e-12, [0.0, 1.0, -3.0, 7.0]]).astype(np.float32)
    got = bin_of(v, bins[1:-1])
    assert got.dtype == np.int8
    assert np.array_equal(got, np.digitize(v, bins[1:-1]))
    for b in range(7):
        assert np.array_equal(np.flatnonzero(got == b), np.flatnonzero((v >= bins[b]) & (v < bins[b + 1])))
