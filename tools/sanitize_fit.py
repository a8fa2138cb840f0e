"""Small fit() for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_fit.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import bench_blobs  # noqa: E402
import annchor_b200.annchor as annchor_mod  # noqa: E402
from annchor_b200.annchor import Annchor  # noqa: E402

if "reorder" in sys.argv[2:]:  # exercise spatial renumbering, tile pruning, scan-ahead and the reduced tile mode at small n
    annchor_mod.REORDER_MIN_POINTS = 0

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
if "hard" in sys.argv[2:]:  # 100 well-separated blobs: empty sampler bins at small n
    X = bench_blobs(n, 128, 100, 42, np.float32)
else:
    X = bench_blobs(n, 16, 40, 2, np.float32)
a = Annchor(X, "euclidean", n_anchors=30, n_neighbors=15, n_samples=2000, p_work=0.1).fit()
print("fit ok", a.evals, a.neighbor_graph[1][:2, :4])
