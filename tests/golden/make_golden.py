#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (gchq/annchor v1.1.0).

Run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py [name ...]

The reference is imported from /root/reference with two import shims
(tests/golden/_shims) for wheels that are absent here (Levenshtein, pynndescent);
everything else -- orchestration, numba kernels, samplers, sklearn regression -- is
the reference's own code.  The stage methods of ``Annchor`` are called in the order
``Annchor.fit`` calls them (annchor/annchor.py:532-623) so intermediate state can be
captured.  Nothing here is read at test time except the .npz outputs.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(HERE, "_shims"), "/root/reference"]

import numpy as np  # noqa: E402


def staged_fit(ann, full):
    """Run ann.fit() stage by stage; return dict of captured arrays."""
    out = {}
    ann.get_anchors()
    out["A"] = np.asarray(ann.A, dtype=np.int64)
    out["D"] = np.ascontiguousarray(ann.D)
    ann.get_locality()
    out["n_pairs"] = np.int64(ann.IJs.shape[0])
    ann.get_features()
    if full:
        out["IJs"] = ann.IJs.astype(np.int32)
        out["features0"] = ann.features.copy()
    for it in range(ann.niters):
        ann.get_sample()
        out["sample_ixs%d" % it] = ann.sample_ixs.astype(np.int64)
        out["sample_ijs%d" % it] = ann.IJs[ann.sample_ixs].astype(np.int32)
        out["sample_bins%d" % it] = ann.sample_bins.copy()
        out["sample_y%d" % it] = ann.sample_y.copy()
        ann.fit_predict_regression()
        out["coef%d" % it] = np.array([lr.coef_ for lr in ann.regression.LRs])
        out["icpt%d" % it] = np.array([lr.intercept_ for lr in ann.regression.LRs])
        if full:
            out["pred%d" % it] = ann.pred.copy()
        ann.fit_predict_errors()
        errs = [ann.error_predictor.errs[b] for b in range(len(ann.error_predictor.errs))]
        out["errs_flat%d" % it] = np.concatenate(errs)
        out["errs_len%d" % it] = np.array([len(e) for e in errs], dtype=np.int64)
        if full:
            out["labels%d" % it] = ann.errors.astype(np.int8)
            out["RA_pre%d" % it] = ann.RefineApprox.copy()
            out["ncm_pre%d" % it] = ann.not_computed_mask.copy()
        ann.select_refine_candidate_pairs(w=1 / ann.niters, it=it)
        out["thresh%d" % it] = ann.thresh.copy()
        out["n_refine%d" % it] = np.int64(len(ann.candidates))
        if full:
            ncm_before = out["ncm_pre%d" % it]
            mapback = np.arange(ncm_before.shape[0])[ncm_before][ann.candidates]
            out["mapback%d" % it] = np.sort(mapback).astype(np.int32)
            out["nextback%d" % it] = np.sort(ann.nextback).astype(np.int32)
        if it < ann.niters - 1:
            ann.update_anchor_points()
            if full:
                out["bounds_upd%d" % it] = ann.features[:, :2].copy()
    ann.get_ann()
    out["ng_idx"] = ann.neighbor_graph[0].astype(np.int32)
    out["ng_dist"] = ann.neighbor_graph[1].copy()
    out["evals"] = np.int64(ann.evals)
    out["p_work"] = np.float64(ann.p_work)
    return out


def blobs(n, d, centers, seed, dtype):
    rng = np.random.default_rng(seed)
    c = rng.normal(size=(centers, d)) * (30.0 / np.sqrt(d))
    lab = rng.integers(0, centers, size=n)
    return (c[lab] + rng.normal(size=(n, d))).astype(dtype)


def save(name, **arrs):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrs)
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def gen_kat():
    """Known answers the reference's own tests assert (tests/test_distances.py:6-12,
    tests/test_datasets.py:107-108,201-202,234-235) plus leaf-kernel vectors."""
    from annchor.distances import levenshtein, euclidean
    from annchor.utils import (get_bounds_njit_ijs, get_dad_ijs, update_bounds, get_probs, get_nn,
                               guarantee_nmin)
    from annchor.datasets import load_strings
    from scipy.spatial.distance import cosine
    from numba import types
    from numba.typed import Dict
    rng = np.random.default_rng(7)
    out = {}
    S = load_strings()
    out["lev_kat_pairs"] = np.array(["cat", "cart", "cat", "cap", "cat", "at", "123456789", "92346781"])
    out["lev_kat_d"] = np.array([levenshtein("cat", "cart"), levenshtein("cat", "cap"),
                                 levenshtein("cat", "at"), levenshtein("123456789", "92346781")])
    assert list(out["lev_kat_d"]) == [1, 1, 1, 3]
    assert levenshtein(S["X"][10], S["X"][165]) == 299
    # leaf kernels on random inputs
    nx, na, n = 300, 12, 4000
    D = rng.random((nx, na)) * 10
    IJ = rng.integers(0, nx, size=(n, 2)).astype(np.int64)
    out["leaf_D"], out["leaf_IJ"] = D, IJ
    out["leaf_bounds"] = get_bounds_njit_ijs(IJ, D)
    out["leaf_dad"] = get_dad_ijs(IJ, D)
    x = rng.random((50, 33)).astype(np.float32)
    out["euc_X"] = x
    out["euc_d"] = np.array([euclidean(x[i], x[j]) for i in range(10) for j in range(10, 20)], dtype=np.float64)
    out["cos_d"] = np.array([cosine(x[i], x[j]) for i in range(10) for j in range(10, 20)], dtype=np.float64)
    # update_bounds on random sparse known lists
    dis = Dict.empty(key_type=types.int64, value_type=types.int64[:])
    ds = Dict.empty(key_type=types.int64, value_type=types.float64[:])
    kptr = [0]
    kids, kds = [], []
    for y in range(nx):
        m = int(rng.integers(0, 40))
        ids = np.sort(rng.choice(nx, size=m, replace=False)).astype(np.int64)
        dd = rng.random(m) * 5
        dis[y], ds[y] = ids, dd
        kids.append(ids)
        kds.append(dd)
        kptr.append(kptr[-1] + m)
    out["ub_kptr"] = np.array(kptr, dtype=np.int64)
    out["ub_kids"] = np.concatenate(kids)
    out["ub_kds"] = np.concatenate(kds)
    out["ub_out"] = update_bounds(IJ, dis, ds)
    # get_probs
    errs = Dict.empty(key_type=types.int64, value_type=types.float64[:])
    flat, ln = [], []
    for b in range(7):
        e = np.sort(rng.normal(size=int(rng.integers(5, 60))))
        errs[b] = e
        flat.append(e)
        ln.append(len(e))
    p = rng.normal(size=n)
    p[:50] = np.concatenate(flat)[:50]  # exact ties with table entries
    lab = rng.integers(0, 7, size=n).astype(np.int64)
    out["pr_p"], out["pr_lab"] = p, lab
    out["pr_errs"], out["pr_len"] = np.concatenate(flat), np.array(ln, dtype=np.int64)
    out["pr_out"] = get_probs(p, np.arange(7), lab, errs)
    save("kat", **out)


def gen_euclid_small():
    """Full stage capture, float64 2-D blobs, n=160."""
    from annchor import Annchor
    from sklearn.datasets import make_blobs
    X, _ = make_blobs(centers=6, n_samples=160, random_state=3)
    ann = Annchor(X, "euclidean", n_anchors=8, n_neighbors=8, n_samples=400, p_work=0.2)
    out = staged_fit(ann, full=True)
    save("euclid_small", X=X, params=np.array([8, 8, 400, 0.2, 42, 2]), **out)


def gen_euclid_f32():
    """Summary capture, float32 d=128 blobs (the bench generator), n=2000."""
    from annchor import Annchor
    X = blobs(2000, 128, 100, 42, np.float32)
    ann = Annchor(X, "euclidean", n_anchors=30, n_neighbors=15, n_samples=2000, p_work=0.1)
    out = staged_fit(ann, full=False)
    save("euclid_f32", params=np.array([30, 15, 2000, 0.1, 42, 2]), gen=np.array([2000, 128, 100, 42]), **out)


def gen_blobs1000():
    """tests/test_examples.py:88-230 config (golden A)."""
    from annchor import Annchor
    from sklearn.datasets import make_blobs
    X, _ = make_blobs(centers=10, n_samples=1000, random_state=42)
    ann = Annchor(X, "euclidean", n_anchors=10, p_work=0.05)
    out = staged_fit(ann, full=False)
    assert list(out["A"]) == [102, 674, 347, 586, 214, 963, 365, 348, 430, 429]
    save("blobs1000", X=X, params=np.array([10, 15, 5000, 0.05, 42, 2]), **out)


def gen_strings():
    """README config (BASELINE config 1) on the bundled strings, plus the bundled exact
    100-NN graph truncated to 30 columns (annchor/data/strings_data.npz)."""
    from annchor import Annchor
    from annchor.datasets import load_strings
    S = load_strings()
    X = S["X"]
    t = time.time()
    ann = Annchor(X, "levenshtein", n_neighbors=25, p_work=0.12)
    out = staged_fit(ann, full=False)
    print("strings fit %.1fs evals %d" % (time.time() - t, out["evals"]))
    text = "\n".join(X.tolist()).encode("ascii")
    save("strings", text=np.frombuffer(text, dtype=np.uint8), y=S["y"].astype(np.int8),
         exact_idx=S["neighbor_graph"][0][:, :30].astype(np.int16),
         exact_dist=S["neighbor_graph"][1][:, :30].astype(np.int16),
         params=np.array([20, 25, 5000, 0.12, 42, 2]), **out)


def gen_w1():
    """1-D Wasserstein on 64-bin histograms, n=400 (small stand-in for BASELINE config 4)."""
    from annchor import Annchor
    rng = np.random.default_rng(5)
    n, nb = 400, 64
    centres = rng.uniform(8, 56, size=(n, 2))
    widths = rng.uniform(1.5, 4.0, size=(n, 2))
    grid = np.arange(nb)[None, :]
    H = sum(np.exp(-0.5 * ((grid - centres[:, k:k + 1]) / widths[:, k:k + 1]) ** 2) for k in range(2))
    H = np.floor(H / H.max(axis=1, keepdims=True) * 255).astype(np.float64)
    M = np.abs(np.arange(nb)[:, None] - np.arange(nb)[None, :]).astype(np.float64)
    ann = Annchor(H, "wasserstein", func_kwargs={"cost_matrix": M}, n_anchors=10, n_neighbors=10,
                  n_samples=700, p_work=0.2)
    out = staged_fit(ann, full=False)
    save("w1", X=H.astype(np.uint8), params=np.array([10, 10, 700, 0.2, 42, 2]), **out)


def gen_cosine():
    """cosine (scipy callable -> joblib path), float64 d=16, n=250."""
    from annchor import Annchor
    rng = np.random.default_rng(11)
    X = blobs(250, 16, 5, 11, np.float64) + 3.0
    ann = Annchor(X, "cosine", n_anchors=8, n_neighbors=10, n_samples=500, p_work=0.2)
    out = staged_fit(ann, full=False)
    save("cosine", X=X, params=np.array([8, 10, 500, 0.2, 42, 2]), **out)


def gen_niters4():
    """The reference's own strings test (annchor/tests/test_annchor.py:71-102): n_anchors=23, k=15,
    p_work=0.12, niters=4 on the bundled strings -- three update_anchor_points rounds, later
    iterations re-fit on what is still not computed.  (The Euclidean blob generators make the
    reference raise "Some sampler bins contain too few samples" in iteration 2 or 3.)"""
    from annchor import Annchor
    from annchor.datasets import load_strings
    X = load_strings()["X"]
    t = time.time()
    ann = Annchor(X, "levenshtein", n_anchors=23, n_neighbors=15, n_samples=5000, p_work=0.12, niters=4)
    out = staged_fit(ann, full=False)
    print("strings niters=4 fit %.1fs evals %d" % (time.time() - t, out["evals"]))
    out.pop("D")  # same strings as strings.npz; anchors differ (23 of them): keep A only
    save("niters4", params=np.array([23, 15, 5000, 0.12, 42, 4]), **out)


def gen_digits():
    """The reference's bundled UCI digits fixture (annchor/data/digits_data.npz, datasets.py:7-46): 1797 8x8
    images as 64-bin histograms, the Euclidean pixel-grid cost matrix, and the first 30 columns of its exact
    100-NN graph -- distances produced by the authors with the real pynndescent kantorovich (general cost
    matrix, exact optimal transport), which is NOT importable here.  Known answers for the device OT kernel and
    for the reference's test_digits / test_brute_force (annchor/tests/test_annchor.py:35-68, 216-245)."""
    from annchor.datasets import load_digits
    d = load_digits()
    X = d["X"]
    assert np.array_equal(X, np.round(X)) and X.min() >= 0 and X.max() <= 255
    ng = d["neighbor_graph"]
    # the sub-graph the reference's test_brute_force builds (tests/test_annchor.py:224-237): first 10 neighbours
    # below 500 of each of the first 500 rows
    small_idx = np.array([ng[0][i][ng[0][i] < 500][:10] for i in range(500)])
    small_dist = np.array([ng[1][i][ng[0][i] < 500][:10] for i in range(500)])
    save("digits", X=X.astype(np.uint8), y=d["y"].astype(np.int8), cost_matrix=d["cost_matrix"].astype(np.float64),
         exact_idx=ng[0][:, :30].astype(np.int16), exact_dist=ng[1][:, :30].astype(np.float64),
         small_idx=small_idx.astype(np.int16), small_dist=small_dist.astype(np.float64))


GENS = {"digits": gen_digits, "niters4": gen_niters4, "kat": gen_kat, "euclid_small": gen_euclid_small, "euclid_f32": gen_euclid_f32,
        "blobs1000": gen_blobs1000, "strings": gen_strings, "w1": gen_w1, "cosine": gen_cosine}

if __name__ == "__main__":
    import annchor  # noqa: F401  (fails loudly if the reference is not importable)
    for nm in (sys.argv[1:] or list(GENS)):
        t0 = time.time()
        try:
            GENS[nm]()
            print("%s done in %.1fs" % (nm, time.time() - t0), flush=True)
        except Exception as e:  # keep going; report at the end
            import traceback
            traceback.print_exc()
            print("%s FAILED: %s" % (nm, e), flush=True)
