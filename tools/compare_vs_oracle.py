"""Device fit() vs the CPU oracle's fit() (= the reference's algorithm) on the same input at a size
the oracle can still run: error counts against the exact graph under the reference's own tie-aware
metric (compare_neighbor_graphs).  The exact graph is computed with the device metric kernel
(all pairs), which the tests pin against the oracle metric.

usage: python tools/compare_vs_oracle.py strings 3000 0.05 25   |   euclid 6000 0.02 15
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import bench_blobs  # noqa: E402
from test_configs_gpu import synthetic_strings  # noqa: E402
from oracle import OracleAnnchor, compare_neighbor_graphs  # noqa: E402
from annchor_b200.annchor import Annchor  # noqa: E402

kind, n, pw, k = sys.argv[1], int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4])
if kind == "strings":
    X, metric = synthetic_strings(n), "levenshtein"
else:
    X, metric = bench_blobs(n, 128, 100, 42, np.float32), "euclidean"
kw = dict(n_anchors=30, n_neighbors=k, n_samples=5000, p_work=pw)
t = time.time()
dev = Annchor(X, metric, **kw).fit()
t_dev = time.time() - t
# exact graph from all-pairs device distances
iu = np.triu_indices(n, 1)
d = dev._dataset.pair_dists(np.stack(iu, axis=1))
Dm = np.zeros((n, n))
Dm[iu] = d
Dm += Dm.T
order = np.argsort(Dm, axis=1, kind="stable")[:, :k]
exact = (order, np.take_along_axis(Dm, order, axis=1))
t = time.time()
orc = OracleAnnchor(X, metric, **kw).fit()
t_orc = time.time() - t
e_dev = compare_neighbor_graphs(exact, dev.neighbor_graph, k)
e_orc = compare_neighbor_graphs(exact, orc.neighbor_graph, k)
print("%s n=%d p_work=%g k=%d: errors device %d, oracle %d (of %d); evals %d / %d; seconds %.2f / %.2f"
      % (kind, n, pw, k, e_dev, e_orc, n * (k - 1), dev.evals, orc.evals, t_dev, t_orc))
