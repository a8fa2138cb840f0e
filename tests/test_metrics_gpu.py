"""K4 / K1 parity: CUDA metric kernels (through the C ABI) vs the CPU oracle and the
reference's golden vectors.  Integer metric: bit-exact.  Float metrics: <= 1e-5 relative
(BASELINE.json north_star); in practice ~1e-7."""
import numpy as np
import pytest

from conftest import load_golden, golden_strings, bench_blobs

pytestmark = pytest.mark.gpu

RTOL = 1e-5  # north_star tolerance for float metrics


def _pairs(n, m, seed=0):
    rng = np.random.default_rng(seed)
    return rng.integers(0, n, size=(m, 2)).astype(np.int64)


def test_levenshtein_kat_and_edge_cases(gpu_ctx):
    import annchor_b200 as ab
    # annchor/tests/test_distances.py:6-12 + empty / ragged strings
    X = ["cat", "cart", "cap", "at", "123456789", "92346781", "", "a" * 70, "a" * 64 + "b" * 64,
         "b" * 129, "abc" * 100]
    ds = ab.Dataset(gpu_ctx, X, "levenshtein")
    IJ = np.array([[0, 1], [0, 2], [0, 3], [4, 5], [6, 0], [0, 6], [6, 6], [7, 8], [8, 9], [9, 10],
                   [10, 7], [3, 3]])
    got = ds.pair_dists(IJ)
    from oracle.metrics import levenshtein
    want = np.array([levenshtein(X[i], X[j]) for i, j in IJ], dtype=np.float64)
    assert list(got[:4]) == [1, 1, 1, 3]
    assert np.array_equal(got, want)


def test_levenshtein_golden_strings(gpu_ctx):
    import annchor_b200 as ab
    from oracle.metrics import PairMetric
    X, g = golden_strings()
    ds = ab.Dataset(gpu_ctx, X, "levenshtein")
    # the bundled exact graph (annchor/data/strings_data.npz): 1600 x 29 golden distances
    rows = np.repeat(np.arange(1600), 29)
    cols = g["exact_idx"][:, 1:30].astype(np.int64).ravel()
    got = ds.pair_dists(np.stack([rows, cols], axis=1))
    assert np.array_equal(got, g["exact_dist"][:, 1:30].astype(np.float64).ravel())
    assert ds.pair_dists(np.array([[10, 165]]))[0] == 299  # tests/test_datasets.py:234-235
    # unsorted random pairs (mixed patterns inside a warp) vs the oracle DP
    IJ = _pairs(1600, 3000, 1)
    assert np.array_equal(ds.pair_dists(IJ), PairMetric(X, "levenshtein")(IJ))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("d", [2, 33, 128, 515])
@pytest.mark.parametrize("metric", ["euclidean", "cosine"])
def test_dense_pairs_vs_oracle(gpu_ctx, dtype, d, metric):
    import annchor_b200 as ab
    from oracle.metrics import PairMetric
    X = bench_blobs(700, d, 7, 3, dtype) + (1.0 if metric == "cosine" else 0.0)
    IJ = _pairs(700, 5001, 2)
    IJ[:5] = [[3, 3], [0, 699], [699, 0], [5, 5], [1, 2]]
    got = ab.Dataset(gpu_ctx, X, metric).pair_dists(IJ)
    want = PairMetric(X, metric)(IJ)
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-6 if metric == "cosine" else 0)
    if metric == "euclidean":
        assert got[0] == 0 and got[3] == 0


def test_dense_kat_vectors(gpu_ctx):
    import annchor_b200 as ab
    k = load_golden("kat")  # values computed by the reference's euclidean / scipy cosine
    ij = np.array([[i, j] for i in range(10) for j in range(10, 20)])
    np.testing.assert_allclose(ab.Dataset(gpu_ctx, k["euc_X"], "euclidean").pair_dists(ij), k["euc_d"],
                               rtol=RTOL)
    np.testing.assert_allclose(ab.Dataset(gpu_ctx, k["euc_X"], "cosine").pair_dists(ij), k["cos_d"],
                               rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("dtype", [np.uint8, np.float64])
def test_w1_pairs(gpu_ctx, dtype):
    import annchor_b200 as ab
    from oracle.metrics import PairMetric
    g = load_golden("w1")
    H = g["X"].astype(dtype)
    IJ = _pairs(H.shape[0], 4000, 4)
    M = np.abs(np.arange(H.shape[1])[:, None] - np.arange(H.shape[1])[None, :]).astype(float)
    got = ab.Dataset(gpu_ctx, H, "wasserstein", cost_matrix=M).pair_dists(IJ)
    np.testing.assert_allclose(got, PairMetric(H.astype(np.float64), "wasserstein1d")(IJ), rtol=1e-9)
    # any other ground cost goes to the exact OT kernel; 2*|a-b| must give exactly twice the closed form's value
    got2 = ab.Dataset(gpu_ctx, H, "wasserstein", cost_matrix=M * 2).pair_dists(IJ[:500])
    np.testing.assert_allclose(got2, 2.0 * got[:500], rtol=1e-9, atol=1e-12)


def test_empty_and_bad_pairs(gpu_ctx):
    import annchor_b200 as ab
    X = bench_blobs(10, 4, 2, 0, np.float32)
    ds = ab.Dataset(gpu_ctx, X, "euclidean")
    assert ds.pair_dists(np.zeros((0, 2), dtype=np.int64)).shape == (0,)
    with pytest.raises(ab.AnnbError):
        ds.pair_dists(np.array([[0, 10]]))


def test_get_exact_ijs_plug(gpu_ctx):
    """the reference's get_exact_ijs(f, X, IJ) contract (annchor/annchor.py:77-82)."""
    import annchor_b200 as ab
    X = bench_blobs(300, 16, 3, 1, np.float64)
    plug = ab.GpuExactIJs("euclidean", gpu_ctx)
    IJ = _pairs(300, 100)
    want = np.array([np.linalg.norm(X[i] - X[j]) for i, j in IJ])
    np.testing.assert_allclose(plug(None, X, IJ), want, rtol=1e-12)


def test_maxmin_golden_A(gpu_ctx):
    """annchor/tests/test_examples.py:228-230: the MaxMin picker's anchors on the blobs data."""
    import annchor_b200 as ab
    g = load_golden("blobs1000")
    ds = ab.Dataset(gpu_ctx, g["X"], "euclidean")
    first = int(np.random.RandomState(42).randint(1000))
    A, D = ds.maxmin_anchors(10, first)
    assert list(A) == [102, 674, 347, 586, 214, 963, 365, 348, 430, 429]
    np.testing.assert_allclose(D, g["D"], rtol=1e-12, atol=1e-12)


def test_maxmin_strings_and_f32(gpu_ctx):
    import annchor_b200 as ab
    X, g = golden_strings()
    ds = ab.Dataset(gpu_ctx, X, "levenshtein")
    A, D = ds.maxmin_anchors(20, int(np.random.RandomState(42).randint(1600)))
    assert np.array_equal(A, g["A"]) and np.array_equal(D, g["D"])
    g2 = load_golden("euclid_f32")
    n, d, c, s = g2["gen"]
    X2 = bench_blobs(int(n), int(d), int(c), int(s), np.float32)
    ds2 = ab.Dataset(gpu_ctx, X2, "euclidean")
    A2, D2 = ds2.maxmin_anchors(30, int(np.random.RandomState(42).randint(int(n))))
    assert np.array_equal(A2, g2["A"])
    np.testing.assert_allclose(D2, g2["D"], rtol=RTOL)
    np.testing.assert_allclose(ds2.anchor_dists(A2), D2, rtol=0, atol=0)
