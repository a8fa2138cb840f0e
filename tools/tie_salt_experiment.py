"""Distribution of the end-to-end error count over tie-break salts (N=2000 reference capture).
Run on the GPU box:  for s in 0 1 2 ...; do ANNB_TIE_SALT=$s python tools/tie_salt_experiment.py; done"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import bench_blobs, load_golden  # noqa: E402
from oracle import OracleBruteForce, compare_neighbor_graphs  # noqa: E402
from annchor_b200.annchor import Annchor  # noqa: E402

g = load_golden("euclid_f32")
n, d, c, s = g["gen"]
X = bench_blobs(int(n), int(d), int(c), int(s), np.float32)
bf = OracleBruteForce(X, "euclidean").fit()
e_ref = compare_neighbor_graphs(bf.neighbor_graph, (g["ng_idx"], g["ng_dist"]), 15)
ann = Annchor(X, "euclidean", n_anchors=30, n_neighbors=15, n_samples=2000, p_work=0.1).fit()
e_dev = compare_neighbor_graphs(bf.neighbor_graph, ann.neighbor_graph, 15)
print("salt", os.environ.get("ANNB_TIE_SALT", "0"), "e_dev", e_dev, "e_ref", e_ref, "evals", ann.evals, int(g["evals"]),
      "tightened", ann.n_tightened)
