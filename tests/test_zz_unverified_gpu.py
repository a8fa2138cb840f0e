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
