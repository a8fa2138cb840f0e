"""Time the device BruteForce (tensor-core path) on the bench workload and report recall of a fit."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_blobs  # noqa: E402
import annchor_b200 as ab  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
X = make_blobs(n, 128, 100, 42)
ctx = ab.default_context()
ds = ab.Dataset(ctx, X, "euclidean")
for it in range(3):
    t = time.time()
    idx, dist = ds.bruteforce_knn(15)
    ctx.sync()
    print("bruteforce_knn N=%d k=15: %.3f s" % (n, time.time() - t), flush=True)
print(idx[:2], dist[:2, :4])
