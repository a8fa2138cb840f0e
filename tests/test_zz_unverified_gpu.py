"""GPU tests written after the round's GPU budget was spent: they have NOT run on hardware yet,
so they are non-strict xfail (an XPASS is the expected outcome) and live in the file pytest runs
last, where a device fault cannot disturb the verified tests.  Promote them to test_index_gpu.py
once they have passed on a B200."""
import numpy as np
import pytest

from conftest import load_golden, golden_strings

pytestmark = pytest.mark.gpu


@pytest.mark.xfail(reason="not yet run on hardware (round 1 GPU budget exhausted)", strict=False)
def test_fit_strings_niters4_reference_test():
    """The reference's own strings test (annchor/tests/test_annchor.py:71-102): n_anchors=23, k=15,
    p_work=0.12, niters=4, `error < 15`.  Three update_anchor_points rounds: the tightening kernel's
    look-up of earlier tightened bounds (has_tight) and repeated TIGHT overwrites are only reached
    with niters > 2.  tests/golden/niters4.npz is the capture of the unmodified reference."""
    from annchor_b200.annchor import Annchor
    from oracle import compare_neighbor_graphs
    from oracle.metrics import PairMetric
    X, gs = golden_strings()
    g = load_golden("niters4")
    ann = Annchor(X, "levenshtein", n_anchors=23, n_neighbors=15, n_samples=5000, p_work=0.12, niters=4).fit()
    assert np.array_equal(ann.A, g["A"])
    assert abs(ann.evals - int(g["evals"])) <= 0.002 * int(g["evals"])
    exact = (gs["exact_idx"].astype(np.int64), gs["exact_dist"].astype(np.float64))
    err = compare_neighbor_graphs(exact, ann.neighbor_graph, 15)
    assert err < 15, err
    idx, dist = ann.neighbor_graph
    ij = np.stack([np.repeat(np.arange(1600), 14), idx[:, 1:].ravel()], axis=1)
    assert np.array_equal(PairMetric(X, "levenshtein")(ij), dist[:, 1:].ravel())


@pytest.mark.xfail(reason="not yet run on hardware (round 1 GPU budget exhausted)", strict=False)
def test_selected_and_random_pickers():
    """annchor/pickers.py:86-128.  SelectedAnchorPicker fed with the anchors MaxMin chose must
    reproduce the MaxMin fit exactly (same D -> same everything); RandomAnchorPicker draws with the
    reference's RandomState rule and must still give a good graph (reference test_examples.py:88-230
    accepts <= 1 error for its custom pickers on this data)."""
    from annchor_b200.annchor import Annchor
    from annchor_b200.plugins import SelectedAnchorPicker, RandomAnchorPicker
    from oracle import OracleBruteForce, compare_neighbor_graphs
    g = load_golden("blobs1000")
    X = g["X"]
    kw = dict(n_anchors=10, p_work=0.05)
    base = Annchor(X, "euclidean", **kw).fit()
    assert np.array_equal(base.A, g["A"])
    sel = Annchor(X, "euclidean", anchor_picker=SelectedAnchorPicker(base.A), **kw).fit()
    assert np.array_equal(sel.A, base.A)
    np.testing.assert_allclose(sel.D, base.D, rtol=1e-12)
    assert np.array_equal(sel.neighbor_graph[0], base.neighbor_graph[0])
    assert np.array_equal(sel.neighbor_graph[1], base.neighbor_graph[1])
    rnd = Annchor(X, "euclidean", anchor_picker=RandomAnchorPicker(), **kw).fit()
    want = np.random.RandomState(42).choice(np.arange(1000), 10, replace=False)
    assert np.array_equal(rnd.A, want)
    exact = OracleBruteForce(X, "euclidean").fit().neighbor_graph
    assert compare_neighbor_graphs(exact, rnd.neighbor_graph, 15) <= 30


@pytest.mark.xfail(reason="not yet run on hardware (round 1 GPU budget exhausted)", strict=False)
def test_is_metric_false_path():
    """is_metric=False (annchor/annchor.py:110,368-372): anchor pairs take their value from D
    instead of relying on lb == ub, the phase-1 lower-bound filters are off, thresholds use the row
    sweep.  On metric data the result must stay a good graph and every distance exact."""
    from annchor_b200.annchor import Annchor
    from oracle import OracleBruteForce, compare_neighbor_graphs
    g = load_golden("blobs1000")
    X = g["X"]
    a = Annchor(X, "euclidean", n_anchors=10, p_work=0.05, is_metric=False).fit()
    exact = OracleBruteForce(X, "euclidean").fit().neighbor_graph
    assert compare_neighbor_graphs(exact, a.neighbor_graph, 15) <= 30
    idx, dist = a.neighbor_graph
    true = np.linalg.norm(X[:, None, :] - X[idx], axis=2)
    np.testing.assert_allclose(dist, true, rtol=1e-5, atol=1e-9)
