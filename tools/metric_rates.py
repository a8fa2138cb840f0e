"""Throughput of the K4 pair-distance kernels (device-resident pairs, CUDA-event timed through the
context timer): Levenshtein in GCUPS (DP cells / s), 1-D Wasserstein and Euclidean in pairs/s and
the HBM-gather rate implied by SURVEY.md 8(d)'s algorithmic bytes."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import annchor_b200 as ab  # noqa: E402
from annchor_b200 import _lib  # noqa: E402
from test_configs_gpu import synthetic_strings, blob_histograms  # noqa: E402
from conftest import bench_blobs  # noqa: E402

ctx = ab.default_context(0)
L = _lib.load()
rng = np.random.default_rng(0)


def rate(ds, n, npairs, reps=5):
    i = torch.from_numpy(rng.integers(0, n, size=npairs).astype(np.int32)).cuda()
    j = torch.from_numpy(rng.integers(0, n, size=npairs).astype(np.int32)).cuda()
    out = torch.empty(npairs, dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        ctx.timer_start()
        _lib.check(L.annb_pair_dists_dev(ctx.handle, ds.handle, ds.metric, i.data_ptr(), j.data_ptr(), npairs,
                                         out.data_ptr()))
        best = min(best, ctx.timer_stop())
    return best * 1e-3, i.cpu().numpy(), j.cpu().numpy()


X = synthetic_strings(20000)
lens = np.array([len(s) for s in X])
ds = ab.Dataset(ctx, X, "levenshtein")
t, i, j = rate(ds, len(X), 4_000_000)
cells = float(np.sum(lens[i].astype(np.float64) * lens[j]))
print("levenshtein len~%d: %.3e pairs/s, %.1f GCUPS" % (lens.mean(), 4e6 / t, cells / t / 1e9))
H = blob_histograms(10000)
ds = ab.Dataset(ctx, H, "wasserstein")
t, _, _ = rate(ds, len(H), 20_000_000)
print("wasserstein1d 784 bins: %.3e pairs/s, %.0f GB/s of CDF rows (2*784*8 B per pair)" % (2e7 / t, 2e7 * 2 * 784 * 8 / t / 1e9))
Xe = bench_blobs(1_000_000, 128, 100, 42, np.float32)
ds = ab.Dataset(ctx, Xe, "euclidean")
t, _, _ = rate(ds, len(Xe), 50_000_000)
print("euclidean d=128 f32 (N=1M rows, random pairs): %.3e pairs/s, %.0f GB/s (1036 B per pair)" % (5e7 / t, 5e7 * 1036 / t / 1e9))
# general cost-matrix Wasserstein (exact OT per pair, one warp per pair): the reference's digits fixture
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
g = load_golden("digits")
ds = ab.Dataset(ctx, g["X"], "wasserstein", cost_matrix=g["cost_matrix"])
t, _, _ = rate(ds, g["X"].shape[0], 2_000_000, reps=3)
print("wasserstein (8x8 digits, general cost matrix, exact OT): %.3e pairs/s" % (2e6 / t))
