"""Device BruteForce (annchor/annchor.py:943-1023; csrc/bruteforce.cu) against the CPU oracle's
exact graph, against the bundled exact fixtures of the reference, and the tensor-core path
against the plain all-pairs path (equality: the GEMM only prunes, the exact metric kernel decides).

Tolerance (float metrics): distances 1e-6 relative vs the float64 oracle (the metric kernels'
own parity is pinned in test_metrics_gpu.py); indices equal wherever distances are not tied."""
import numpy as np
import pytest

from conftest import load_golden, golden_strings, bench_blobs

pytestmark = pytest.mark.gpu


def _oracle_graph(X, metric, k):
    from oracle import OracleBruteForce
    g = OracleBruteForce(X, metric).fit().neighbor_graph
    return g[0][:, :k], g[1][:, :k]


def _check_vs_oracle(dev, orc, k, rtol):
    from oracle import compare_neighbor_graphs
    assert dev[0].shape == orc[0].shape == (dev[0].shape[0], k)
    assert np.array_equal(dev[0][:, 0], np.arange(dev[0].shape[0])) and np.all(dev[1][:, 0] == 0)
    np.testing.assert_allclose(dev[1], orc[1], rtol=rtol, atol=1e-9)
    assert compare_neighbor_graphs(orc, dev, k) == 0
    untied = np.ones(orc[1].shape, bool)
    untied[:, 1:] &= np.abs(np.diff(orc[1], axis=1)) > 1e-6 * (1 + orc[1][:, 1:])
    untied[:, :-1] &= np.abs(np.diff(orc[1], axis=1)) > 1e-6 * (1 + orc[1][:, 1:])
    assert np.array_equal(dev[0][untied], orc[0][untied])


@pytest.mark.parametrize("n,d,metric", [(3000, 128, "euclidean"), (2500, 128, "cosine"), (1111, 20, "euclidean"),
                                        (4000, 70, "cosine")])
def test_tensor_core_path_vs_oracle(gpu_ctx, n, d, metric):
    from annchor_b200.annchor import BruteForce
    X = bench_blobs(n, d, 25, 3, np.float32)
    k = 15
    bf = BruteForce(X, metric, n_neighbors=k, ctx=gpu_ctx).fit()
    _check_vs_oracle(bf.neighbor_graph, _oracle_graph(X, metric, k), k, rtol=2e-6)


@pytest.mark.parametrize("metric,dtype", [("euclidean", np.float32), ("cosine", np.float32), ("euclidean", np.float64)])
def test_tensor_core_path_equals_all_pairs_path(gpu_ctx, monkeypatch, metric, dtype):
    """Same exact metric kernel decides in both paths -> identical arrays, ties included.  Heavy
    ties on purpose: integer-valued rows (many equal distances) and duplicated points."""
    from annchor_b200.annchor import BruteForce
    rng = np.random.default_rng(4)
    X = np.concatenate([bench_blobs(5000, 128, 40, 5, dtype), rng.integers(0, 4, size=(1500, 128)).astype(dtype)])
    X[100:110] = X[0]  # duplicates: distance exactly 0
    k = 25
    a = BruteForce(X, metric, n_neighbors=k, ctx=gpu_ctx).fit().neighbor_graph
    monkeypatch.setenv("ANNB_BRUTEFORCE_EXACT", "1")
    b = BruteForce(X, metric, n_neighbors=k, ctx=gpu_ctx).fit().neighbor_graph
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[1], b[1])


def test_strings_equal_bundled_exact_graph(gpu_ctx):
    """The reference's bundled exact 100-NN graph of load_strings() (annchor/data/strings_data.npz;
    first 30 columns in tests/golden/strings.npz): error == 0 and bit-exact distances."""
    from annchor_b200.annchor import BruteForce
    from oracle import compare_neighbor_graphs
    X, g = golden_strings()
    k = 30
    bf = BruteForce(X, "levenshtein", n_neighbors=k, ctx=gpu_ctx).fit()
    exact = (g["exact_idx"].astype(np.int64), g["exact_dist"].astype(np.float64))
    assert compare_neighbor_graphs(exact, bf.neighbor_graph, k) == 0
    assert np.array_equal(bf.neighbor_graph[1], exact[1])


def test_reference_signature_small(gpu_ctx):
    """annchor/tests/test_annchor.py:216-248 shape: BruteForce(X, metric).fit() returns every column
    for small nx; 1-D Wasserstein histograms go through the all-pairs path."""
    from annchor_b200.annchor import BruteForce
    g = load_golden("w1")
    H = g["X"]
    M = np.abs(np.arange(H.shape[1])[:, None] - np.arange(H.shape[1])[None, :]).astype(float)
    bf = BruteForce(H, "wasserstein", func_kwargs={"cost_matrix": M}, ctx=gpu_ctx).fit()
    assert bf.neighbor_graph[0].shape == (400, 400)
    orc = _oracle_graph(H.astype(np.float64), "wasserstein1d", 400)
    np.testing.assert_allclose(bf.neighbor_graph[1], orc[1], rtol=1e-9, atol=1e-12)
    with pytest.raises(ValueError):
        BruteForce(np.zeros((5000, 3), np.float32), "euclidean", ctx=gpu_ctx)
    with pytest.raises(NotImplementedError):
        BruteForce(H, lambda a, b: 0.0)
