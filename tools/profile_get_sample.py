"""cProfile of the host side of get_sample (N given on the command line)."""
import cProfile, pstats, sys
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from bench import make_blobs
import annchor_b200 as ab
from annchor_b200.annchor import Annchor
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
X = make_blobs(n, 128, 100, 42)
ctx = ab.default_context()
for it in range(2):
    a = Annchor(X, "euclidean", ctx=ctx, n_anchors=30, n_neighbors=15, n_samples=5000, p_work=0.001 if n >= 500000 else 0.01)
    a.get_anchors()  # (renumbers the points itself at large n)
    a.get_locality()
    pr = cProfile.Profile()
    pr.enable(); a.get_sample(); ctx.sync(); pr.disable()
    if it == 1:
        pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
    a._index.close()
