"""General cost-matrix Wasserstein (annchor/utils.py:75-86: kantorovich(x, y, cost=M)) on the device: exact optimal
transport per pair (ot_pair_kernel), checked against the reference's own bundled digits fixture -- the distances of its
exact 100-NN graph were produced by the authors with the real pynndescent kantorovich -- and the reference's tests on
it (annchor/tests/test_annchor.py:35-68 test_digits, 216-245 test_brute_force)."""
import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


def _digits():
    g = load_golden("digits")
    return g["X"], g["cost_matrix"], (g["exact_idx"].astype(np.int64), g["exact_dist"])


def test_ot_kernel_equals_reference_distances(gpu_ctx):
    import annchor_b200 as ab
    X, M, (idx, dist) = _digits()
    n = X.shape[0]
    ds = ab.Dataset(gpu_ctx, X, "wasserstein", cost_matrix=M)
    assert ds.metric == ab._lib.WASSERSTEIN
    IJ = np.stack([np.repeat(np.arange(n), idx.shape[1]), idx.ravel()], 1)
    got = ds.pair_dists(IJ)
    np.testing.assert_allclose(got, dist.ravel(), rtol=1e-9, atol=1e-12)
    # symmetric; float64 input and scaled masses give the same values (unit-mass normalisation)
    np.testing.assert_allclose(ds.pair_dists(IJ[:3000, ::-1]), got[:3000], rtol=1e-12, atol=1e-14)
    ds64 = ab.Dataset(gpu_ctx, X.astype(np.float64) * 3.5, "wasserstein", cost_matrix=M)
    np.testing.assert_allclose(ds64.pair_dists(IJ[:3000]), got[:3000], rtol=1e-12, atol=1e-14)
    # the anchor-row entry (one item against all) is the same kernel
    a = np.array([5, 900], dtype=np.int64)
    D = ds.anchor_dists(a)
    for k, ai in enumerate(a):
        pairs = np.stack([np.full(200, ai), np.arange(200)], 1)
        assert np.array_equal(D[:200, k], ds.pair_dists(pairs))


def test_ot_kernel_equals_oracle_on_random_pairs(gpu_ctx):
    import annchor_b200 as ab
    from oracle.metrics import PairMetric
    X, M, _ = _digits()
    rng = np.random.default_rng(3)
    IJ = rng.integers(0, X.shape[0], size=(6000, 2))
    got = ab.Dataset(gpu_ctx, X, "wasserstein", cost_matrix=M).pair_dists(IJ)
    want = PairMetric(X.astype(np.float64), "wasserstein", cost_matrix=M)(IJ)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)
    # a non-symmetric, non-metric cost matrix and sparse histograms (ragged supports, single-bin masses)
    nb = 37
    C = rng.uniform(0.0, 5.0, size=(nb, nb))
    H = rng.integers(0, 4, size=(300, nb)).astype(np.float64) * (rng.random((300, nb)) < 0.3)
    H[H.sum(1) == 0, 0] = 1.0
    H[:5] = 0.0
    H[np.arange(5), np.arange(5)] = 2.0  # point masses: distance = the cost entry
    IJ = rng.integers(0, 300, size=(4000, 2))
    got = ab.Dataset(gpu_ctx, H, "wasserstein", cost_matrix=C).pair_dists(IJ)
    want = PairMetric(H, "wasserstein", cost_matrix=C)(IJ)
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)
    pm = ab.Dataset(gpu_ctx, H, "wasserstein", cost_matrix=C).pair_dists(np.array([[0, 1], [3, 2]]))
    np.testing.assert_allclose(pm, [C[0, 1], C[3, 2]], rtol=1e-12)
    with pytest.raises(ab.AnnbError):
        ab.Dataset(gpu_ctx, np.ones((4, 65)), "wasserstein", cost_matrix=np.ones((65, 65)))


def test_digits_fit_as_the_reference_tests_it(gpu_ctx):
    """annchor/tests/test_annchor.py:35-68: k=25, n_anchors=25, n_samples=5000, p_work=0.16 -> fewer than 10 errors
    against the bundled exact graph."""
    from annchor_b200.annchor import Annchor
    from oracle import compare_neighbor_graphs
    X, M, exact = _digits()
    k = 25
    ann = Annchor(X, "wasserstein", func_kwargs={"cost_matrix": M}, n_anchors=25, n_neighbors=k, n_samples=5000,
                  p_work=0.16, random_seed=42, ctx=gpu_ctx).fit()
    err = compare_neighbor_graphs(exact, ann.neighbor_graph, k)
    assert err < 10, err


def test_digits_bruteforce_as_the_reference_tests_it(gpu_ctx):
    """annchor/tests/test_annchor.py:216-245: BruteForce on the first 500 digits equals the exact graph restricted
    to them (10 neighbours, zero errors)."""
    from annchor_b200.annchor import BruteForce
    from oracle import compare_neighbor_graphs
    X, M, _ = _digits()
    g = load_golden("digits")
    small = (g["small_idx"].astype(np.int64), g["small_dist"])
    bf = BruteForce(X[:500], "wasserstein", func_kwargs={"cost_matrix": M}, ctx=gpu_ctx).fit()
    assert compare_neighbor_graphs(small, bf.neighbor_graph, 10) == 0
