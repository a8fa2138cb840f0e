"""Streaming-index parity (K1..K5 through the C ABI) against the CPU oracle, stage by stage on
identical state, and end to end under the reference's own tie-aware metric
(compare_neighbor_graphs, annchor/annchor.py:1026-1066).

Tolerances: the sweeps compute bounds / predictions in float32 from a float32 copy of D, the
oracle (like the reference) in float64 -- values agree to ~1e-6 relative; sets selected by a
threshold on those values agree except for pairs within rounding distance of the cut."""
import numpy as np
import pytest

from conftest import load_golden, golden_strings, bench_blobs

pytestmark = pytest.mark.gpu


def _oracle_at_select(X, metric, na, nn, ns, pw, seed=42):
    """Oracle advanced to just before select_refine_candidate_pairs of iteration 0."""
    import oracle.pipeline as P
    from oracle import OracleAnnchor
    o = OracleAnnchor(X, metric, n_anchors=na, n_neighbors=nn, n_samples=ns, p_work=pw, random_seed=seed)
    o.get_anchors()
    o.get_locality()
    o.get_features()
    o.get_sample()
    o.fit_predict_regression()
    o.fit_predict_errors()
    return o


def _index_from_oracle(gpu_ctx, o, X, metric):
    import annchor_b200 as ab
    from annchor_b200.annchor import Index
    ds = ab.Dataset(gpu_ctx, X, metric)
    ix = Index(gpu_ctx, ds, o.n_anchors, o.n_neighbors, o.locality, o.loc_thresh, o.loc_min)
    ix.set_anchors(o.A, o.D)
    ncand, nrelax = ix.locality()
    assert ncand == o.IJs.shape[0]
    ix.add_known(o.IJs[o.sample_ixs], o.sample_y)
    eptr = np.zeros(len(o.errs) + 1, dtype=np.int64)
    np.cumsum([len(e) for e in o.errs], out=eptr[1:])
    ix.set_model(o.sample_bins, o.coef, o.icpt, np.concatenate(o.errs), eptr)
    return ds, ix


CASES = {
    "euclid_small": lambda: (load_golden("euclid_small")["X"], "euclidean", 8, 8, 400, 0.2),
    "blobs1000": lambda: (load_golden("blobs1000")["X"], "euclidean", 10, 15, 5000, 0.05),
    "f32_d128": lambda: (bench_blobs(2000, 128, 100, 42, np.float32), "euclidean", 30, 15, 2000, 0.1),
    "strings": lambda: (golden_strings()[0], "levenshtein", 20, 25, 5000, 0.12),
}


@pytest.mark.parametrize("case", list(CASES))
def test_anchors_locality_thresh_select(gpu_ctx, case):
    import oracle.pipeline as P
    X, metric, na, nn, ns, pw = CASES[case]()
    o = _oracle_at_select(X, metric, na, nn, ns, pw)
    ds, ix = _index_from_oracle(gpu_ctx, o, X, metric)

    # --- K1: the device MaxMin picker reproduces the oracle's anchors and D
    import annchor_b200 as ab
    first = int(np.random.RandomState(42).randint(len(X)))
    A, D = ds.maxmin_anchors(na, first)
    assert np.array_equal(A, o.A)
    np.testing.assert_allclose(D, o.D, rtol=1e-5, atol=1e-9)

    # --- sampler support: the exact pool is precisely the not-computed candidate set
    ncm = o.not_computed_mask
    n_pool, n_nc, exact = ix.sample_pool(7, 4_000_000)
    assert exact and n_pool == n_nc == int(ncm.sum())
    pij, pdad = ix.get_pool()
    order = np.argsort(pij[:, 0] * len(X) + pij[:, 1])
    assert np.array_equal(pij[order], o.IJs[ncm])
    np.testing.assert_allclose(pdad[order], o.features[ncm][:, 2], rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(ix.pair_features(o.IJs[:500]), o.features[:500, :3], rtol=2e-6, atol=1e-6)

    # --- K5a: thresh (annchor.py:399-404)
    th_o = P.row_kth(o.RefineApprox, o.row_ptr, o.row_pairs, nn)
    th_d = ix.row_thresh()
    # float32 sweep vs float64 oracle: equal to rounding, except rows whose k-th pair is the pair
    # sitting exactly ON a regression-bin edge (q1 / q3 are dad values of actual pairs; the oracle's
    # float64 edge and the device's float32 dad can land on opposite sides of `dad > edge`)
    close = np.isclose(th_d, th_o, rtol=2e-5, atol=1e-5)
    assert close.mean() >= 0.98, (np.nonzero(~close)[0], th_d[~close], th_o[~close])

    # --- guarantee_nmin + scoring + selection (annchor.py:407-465): counts and structure here; the
    # selected / look-ahead SETS, n_forced and the thresholds of both iterations are compared for
    # equality with the device-arithmetic oracle in tests/test_exact_parity_gpu.py, and
    # test_select_is_top_by_probability below checks the cut against the float64 probabilities
    nc_before = o.IJs[ncm]  # (the oracle's selection below clears the mask of what it picks)
    o.select_refine_candidate_pairs(0.5, 0)
    ix.guarantee_nmin(3 * nn // 2)
    n_sel, n_next = ix.select(o.n_refine, o.lookahead)
    sel, nxt = ix.get_selected()
    assert n_sel == min(o.n_refine, o.prob.shape[0]) == sel.shape[0]
    key = lambda ij: set((ij[:, 0] * len(X) + ij[:, 1]).tolist())
    s_dev = key(sel)
    assert len(s_dev) == n_sel and np.all(sel[:, 0] < sel[:, 1])
    assert s_dev <= key(nc_before), "selected pairs must be not-computed candidates"
    assert not (key(nxt) & s_dev)
    if o.n_refine * o.lookahead < o.prob.shape[0]:
        assert nxt.shape[0] == o.n_refine * (o.lookahead - 1)


def test_select_is_top_by_probability(gpu_ctx):
    """Every selected pair's oracle probability >= every unselected pair's (up to the float32
    level-flip tolerance), i.e. the histogram cut equals np.argpartition's."""
    X, metric, na, nn, ns, pw = CASES["blobs1000"]()
    o = _oracle_at_select(X, metric, na, nn, ns, pw)
    ds, ix = _index_from_oracle(gpu_ctx, o, X, metric)
    ix.guarantee_nmin(3 * nn // 2)
    ncm = o.not_computed_mask.copy()
    o.select_refine_candidate_pairs(0.5, 0)
    ix.select(o.n_refine, o.lookahead)
    sel, _ = ix.get_selected()
    # oracle prob per not-computed pair
    ij_nc = o.IJs[ncm]
    lut = {int(a) * len(X) + int(b): p for (a, b), p in zip(ij_nc, o.prob)}
    p_sel = np.array([lut[int(a) * len(X) + int(b)] for a, b in sel])
    cut = np.sort(o.prob)[::-1][o.n_refine - 1]
    # at most a handful of picks may sit one level below the cut (float32 vs float64 rank flips)
    assert np.mean(p_sel >= cut - 2.0 / 600) > 0.995
    assert np.mean(p_sel >= cut) > 0.97


def _fit_dev(X, metric, cost=None, **kw):
    import annchor_b200 as ab
    from annchor_b200.annchor import Annchor
    fk = {"cost_matrix": cost} if cost is not None else None
    return Annchor(X, metric, func_kwargs=fk, **kw).fit()


def test_fit_blobs1000_reference_test(gpu_ctx):
    """annchor/tests/test_examples.py:88-112: MaxMin anchors equal the golden A and the graph has
    zero errors against brute force."""
    from oracle import OracleBruteForce, compare_neighbor_graphs
    g = load_golden("blobs1000")
    ann = _fit_dev(g["X"], "euclidean", n_anchors=10, p_work=0.05)
    assert list(ann.A) == [102, 674, 347, 586, 214, 963, 365, 348, 430, 429]
    assert ann.evals == int(g["evals"])
    bf = OracleBruteForce(g["X"], "euclidean").fit()
    assert compare_neighbor_graphs(bf.neighbor_graph, ann.neighbor_graph, 15) == 0
    assert np.array_equal(ann.neighbor_graph[0][:, 0], np.arange(1000))
    assert np.all(ann.neighbor_graph[1][:, 0] == 0)


def test_fit_strings_readme_config(gpu_ctx):
    """BASELINE config 1 (README.md:102): load_strings, levenshtein, k=25, p_work=0.12.
    Integer metric: every emitted distance is an exact metric value -> rows equal the bundled
    exact graph as sorted multisets wherever correct; reference had 0 errors, test bound is < 15."""
    from oracle import compare_neighbor_graphs
    X, g = golden_strings()
    ann = _fit_dev(X, "levenshtein", n_neighbors=25, p_work=0.12)
    assert np.array_equal(ann.A, g["A"])
    assert np.array_equal(ann.D, g["D"])
    assert ann.evals == int(g["evals"])
    exact = (g["exact_idx"].astype(np.int64), g["exact_dist"].astype(np.float64))
    err = compare_neighbor_graphs(exact, ann.neighbor_graph, 25)
    assert err < 15, err
    # bit-exactness of what is emitted: each (i, j, d) is the true Levenshtein distance
    from oracle.metrics import PairMetric
    idx, dist = ann.neighbor_graph
    rows = np.repeat(np.arange(1600), 24)
    ij = np.stack([rows, idx[:, 1:].ravel()], axis=1)
    assert np.array_equal(PairMetric(X, "levenshtein")(ij), dist[:, 1:].ravel())


def test_fit_f32_d128_vs_reference_capture(gpu_ctx):
    """The bench generator at N=2000: error count vs exact brute force no worse than the
    reference's own fit() on the same input (tests/golden/euclid_f32.npz)."""
    from oracle import OracleBruteForce, compare_neighbor_graphs
    g = load_golden("euclid_f32")
    n, d, c, s = g["gen"]
    X = bench_blobs(int(n), int(d), int(c), int(s), np.float32)
    ann = _fit_dev(X, "euclidean", n_anchors=30, n_neighbors=15, n_samples=2000, p_work=0.1)
    assert np.array_equal(ann.A, g["A"])
    # iteration 1's lowest dad bin is nearly exhausted (the reference drew 1770 of 2000 samples);
    # how many pairs are left there depends on tie-breaks of the previous selection
    assert abs(ann.evals - int(g["evals"])) <= 0.001 * int(g["evals"])
    bf = OracleBruteForce(X, "euclidean").fit()
    e_ref = compare_neighbor_graphs(bf.neighbor_graph, (g["ng_idx"], g["ng_dist"]), 15)
    e_dev = compare_neighbor_graphs(bf.neighbor_graph, ann.neighbor_graph, 15)
    # The count depends on how ties at the selection cuts fall (prob has <= ~716 levels per label):
    # the oracle gives 524 with numpy's argpartition order (= the reference), 605 with a stable
    # order, 552..585 with random tie-breaks; the device breaks ties by a salted hash of the pair and
    # gives 562..702 over 8 salts (tools/tie_salt_experiment.py).
    assert e_dev <= 1.4 * e_ref + 30, (e_dev, e_ref)
    # every emitted distance is a true distance (float32 metric, 1e-5 relative)
    idx, dist = ann.neighbor_graph
    true = np.linalg.norm(X[:, None, :].astype(np.float64)[:200] - X[idx[:200]].astype(np.float64), axis=2)
    np.testing.assert_allclose(dist[:200], true, rtol=1e-5, atol=1e-6)


def test_fit_w1_and_cosine(gpu_ctx):
    from oracle import OracleBruteForce, compare_neighbor_graphs
    g = load_golden("w1")
    H = g["X"]
    M = np.abs(np.arange(H.shape[1])[:, None] - np.arange(H.shape[1])[None, :]).astype(float)
    ann = _fit_dev(H, "wasserstein", cost=M, n_anchors=10, n_neighbors=10, n_samples=700, p_work=0.2)
    assert np.array_equal(ann.A, g["A"])
    bf = OracleBruteForce(H.astype(np.float64), "wasserstein1d").fit()
    e_ref = compare_neighbor_graphs(bf.neighbor_graph, (g["ng_idx"], g["ng_dist"]), 10)
    assert compare_neighbor_graphs(bf.neighbor_graph, ann.neighbor_graph, 10) <= 1.1 * e_ref + 10
    g = load_golden("cosine")
    ann = _fit_dev(g["X"], "cosine", n_anchors=8, n_neighbors=10, n_samples=500, p_work=0.2)
    assert np.array_equal(ann.A, g["A"])
    bf = OracleBruteForce(g["X"], "cosine").fit()
    e_ref = compare_neighbor_graphs(bf.neighbor_graph, (g["ng_idx"], g["ng_dist"]), 10)
    # cosine is not a metric and n=250 is tiny: the error count is noisy (reference: 273)
    assert compare_neighbor_graphs(bf.neighbor_graph, ann.neighbor_graph, 10) <= 1.1 * e_ref + 10


def test_sampler_and_regression_equal_reference_capture(gpu_ctx):
    """Exact-mode sampler: same pairs, bins, distances and regression coefficients as the
    unmodified reference drew on the same input (tests/golden/blobs1000.npz, strings.npz) --
    the numba MT19937 Fisher-Yates stream of annchor/utils.py:555-557,572 is reproduced."""
    from annchor_b200.annchor import Annchor
    gs = load_golden("euclid_small")
    for X_, g_, kw in ((load_golden("blobs1000")["X"], load_golden("blobs1000"),
                        dict(n_anchors=10, p_work=0.05)),
                       (gs["X"], gs, dict(n_anchors=8, n_neighbors=8, n_samples=400, p_work=0.2))):
        ann = Annchor(X_, "euclidean", **kw)
        ann.get_anchors()
        ann.get_locality()
        assert ann.n_candidates == int(g_["n_pairs"])
        ann.get_sample()
        assert np.array_equal(ann.sample_ijs, g_["sample_ijs0"])
        np.testing.assert_allclose(ann.sample_bins, g_["sample_bins0"], rtol=1e-12)
        np.testing.assert_allclose(ann.sample_y, g_["sample_y0"], rtol=1e-6)
        ann.fit_predict_regression()
        coef = np.array([lr.coef_ for lr in ann.regression.LRs])
        np.testing.assert_allclose(coef, g_["coef0"], rtol=1e-4, atol=1e-6)
    # integer metric: ties among equidistant anchors are ordered by np.argsort's unstable sort in
    # the reference (annchor.py:235) and by anchor index here -> the candidate sets differ by a
    # handful of pairs (23 of 1 279 175 on the bundled strings)
    X, g = golden_strings()
    ann = Annchor(X, "levenshtein", n_neighbors=25, p_work=0.12)
    ann.get_anchors()
    ann.get_locality()
    assert abs(ann.n_candidates - int(g["n_pairs"])) <= 1e-4 * int(g["n_pairs"])


def test_p_work_clamps_and_errors(gpu_ctx):
    from annchor_b200.annchor import Annchor
    X = np.random.default_rng(0).random((100, 3))
    assert Annchor(X, "euclidean", p_work=1.5).p_work == 1.0
    a = Annchor(X, "euclidean", p_work=0.0)
    assert a.p_work == min((2 * (a.na + a.n_samples) + 1) / a.N, 1)  # tests/test_annchor.py:148-160
    with pytest.raises(NotImplementedError):
        Annchor(X, lambda x, y: 0.0)
    with pytest.raises(AssertionError):
        Annchor(X, "manhattan")


@pytest.mark.parametrize("kind", ["euclid", "strings"])
def test_two_stage_thresholds_equal_row_sweep(gpu_ctx, monkeypatch, kind):
    """Large metric problems compute thresh / the guarantee_nmin lists in two stages (column-subset
    bounds, one visit per pair, per-row selection; sweep_thresh.cu).  The result must equal the row
    sweep's bit for bit: thresholds directly, and the whole fit() (which also exercises the
    guarantee_nmin lists of iteration 0) through the final graph."""
    from annchor_b200.annchor import Annchor
    if kind == "euclid":
        X, metric = bench_blobs(20000, 128, 100, 42, np.float32), "euclidean"
    else:  # integer metric: massive ties at every cut
        from test_configs_gpu import synthetic_strings
        X, metric = synthetic_strings(20000), "levenshtein"
    kw = dict(n_anchors=30, n_neighbors=15, n_samples=5000, p_work=0.01)

    def staged():
        a = Annchor(X, metric, ctx=gpu_ctx, **kw)
        a.get_anchors()
        a.get_locality()
        a.get_sample()
        a.fit_predict_regression()
        a.fit_predict_errors()
        return a

    a = staged()
    monkeypatch.delenv("ANNB_THRESH_ROWS", raising=False)
    th_two = a._index.row_thresh()
    monkeypatch.setenv("ANNB_THRESH_ROWS", "1")
    th_rows = a._index.row_thresh()
    assert np.array_equal(th_two, th_rows)
    assert np.isfinite(th_rows).all()
    full_rows = Annchor(X, metric, ctx=gpu_ctx, **kw).fit()
    monkeypatch.delenv("ANNB_THRESH_ROWS")
    full_two = Annchor(X, metric, ctx=gpu_ctx, **kw).fit()
    assert full_two.n_forced == full_rows.n_forced
    assert full_two.evals == full_rows.evals
    assert np.array_equal(full_two.neighbor_graph[0], full_rows.neighbor_graph[0])
    assert np.array_equal(full_two.neighbor_graph[1], full_rows.neighbor_graph[1])


# ---- promoted from the round-1 "unverified" file: all three passed on the driver's B200 (GPUTEST_r01) ----
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
