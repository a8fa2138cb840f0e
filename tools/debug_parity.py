"""Debug helper: run one exact-parity case and print what differs between the device fit and the
device-arithmetic oracle (tests/test_exact_parity_gpu.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_exact_parity_gpu as T  # noqa: E402
import annchor_b200 as ab  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "f32_d128"
X, metric, kw, cost = T.CASES[case]()
dev, orc, td, to = T._run_both(ab.default_context(), X, metric, kw, cost)
n = len(X)
for it in range(kw.get("niters", 2)):
    for nm in ("sample_ijs", "sample_bins", "thresh"):
        a, b = td["%s%d" % (nm, it)], to["%s%d" % (nm, it)]
        print(it, nm, "equal" if np.array_equal(a, b) else "DIFF %d" % int(np.sum(a != b)))
    for nm in ("selected", "next"):
        a, b = T._pairset(td["%s%d" % (nm, it)], n), T._pairset(to["%s%d" % (nm, it)], n)
        od, oo = np.setdiff1d(a, b), np.setdiff1d(b, a)
        print(it, nm, a.size, "only dev", od.size, "only oracle", oo.size)
    if "n_tightened%d" % it in td:
        print(it, "n_tightened", td["n_tightened%d" % it], to["n_tightened%d" % it])
print("n_forced", td["n_forced"], to["n_forced"], "evals", dev.evals, orc.evals)
print("graph idx equal", np.array_equal(dev.neighbor_graph[0], orc.neighbor_graph[0]),
      "dist equal", np.array_equal(dev.neighbor_graph[1], orc.neighbor_graph[1]))
