"""get_nearest_enemies / alpha_rss (annchor/annchor.py:685-786, 921-940) on the device: the nearest-enemy graph is exact
here (the reference approximates it from its materialised state), alpha_rss makes the reference's decisions."""
import numpy as np
import pytest

from conftest import bench_blobs, load_golden

pytestmark = pytest.mark.gpu


def _exact_enemies(Dm, y, nn):
    n = len(y)
    idx = np.empty((n, nn), dtype=np.int64)
    dist = np.empty((n, nn))
    for i in range(n):
        cand = np.nonzero(y != y[i])[0]
        o = np.lexsort((cand, Dm[i, cand]))[:nn]
        idx[i], dist[i] = cand[o], Dm[i, cand[o]]
    return idx, dist


def _reference_alpha_rss(Dm, dne, alpha):
    """annchor/annchor.py:921-940 restated on a full distance matrix."""
    ix = np.argsort(dne)
    rss = [ix[0]]
    adne = dne / (1 + alpha)
    for i in ix:
        dnn = np.min(Dm[i, rss])
        if (dnn > adne[i]) or np.isclose(dnn, adne[i]):
            rss.append(i)
    return np.array(rss)


@pytest.mark.parametrize("metric", ["euclidean", "wasserstein"])
def test_nearest_enemies_exact_and_alpha_rss(gpu_ctx, metric):
    from annchor_b200.annchor import Annchor
    import annchor_b200 as ab
    if metric == "euclidean":
        rng = np.random.default_rng(5)
        n = 1500
        X = bench_blobs(n, 16, 40, 2, np.float32)
        y = rng.integers(0, 4, size=n)
        y[:7] = 9  # a small class
        kw, fk = dict(n_anchors=15, n_neighbors=10, n_samples=1000, p_work=0.1), None
    else:
        g = load_golden("digits")
        n = 600
        X, y = g["X"][:n], g["y"][:n].astype(np.int64)
        kw, fk = dict(n_anchors=15, n_neighbors=10, n_samples=1000, p_work=0.16), {"cost_matrix": g["cost_matrix"]}
    ann = Annchor(X, metric, func_kwargs=fk, ctx=gpu_ctx, **kw).fit()
    I, J = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    Dm = ann._dataset.pair_dists(np.stack([I.ravel(), J.ravel()], 1)).reshape(n, n)
    for nn in (1, 3):
        ngi, ngd = ann.get_nearest_enemies(y, nn=nn)
        wi, wd = _exact_enemies(Dm, y, nn)
        assert np.array_equal(ngd, wd)
        assert np.array_equal(ngi, wi)
        assert (y[ngi] != y[:, None]).all()
    dne = ann.nearest_enemy_graph[1][:, 0]
    for alpha in (0, 0.2):
        got = ann.alpha_rss(y, alpha=alpha)
        want = _reference_alpha_rss(Dm, dne, alpha)
        assert np.array_equal(got, want)
        # the defining property: every point has a kept point within its (relaxed) nearest-enemy distance
        assert (Dm[:, got].min(1) <= dne / (1 + alpha) + 1e-12).all()
    with pytest.raises(AssertionError):
        ann.get_nearest_enemies(np.zeros(n), nn=1)          # one label only
    with pytest.raises(AssertionError):
        ann.get_nearest_enemies(y[:-1], nn=1)               # length mismatch
    if metric == "euclidean":
        with pytest.raises(AssertionError):
            ann.get_nearest_enemies(y, nn=8)                # the small class has 7 members
    with pytest.raises(ab.AnnbError):
        ann._dataset.nearest_enemies(np.arange(n) % 2, n)   # nn out of range
