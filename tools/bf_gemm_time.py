"""Time the BruteForce GEMM kernel alone under the ANNB_BF_DEBUG variants (timing experiments)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import time
    from bench import make_blobs
    import annchor_b200 as ab
    n = int(sys.argv[2])
    X = make_blobs(n, 128, 100, 42)
    ctx = ab.default_context()
    ds = ab.Dataset(ctx, X, "euclidean")
    for it in range(2):
        t = time.time()
        try:
            ds.bruteforce_knn(15)
        except Exception as e:
            print("  (result rejected: %s)" % str(e)[:80])
        ctx.sync()
    print("debug=%s N=%d: %.3f s" % (os.environ.get("ANNB_BF_DEBUG", "0"), n, time.time() - t), flush=True)
else:
    n = sys.argv[1] if len(sys.argv) > 1 else "100000"
    for dbg in ("0", "1", "2", "3"):
        env = dict(os.environ, ANNB_BF_DEBUG=dbg, ANNB_TRACE="1")
        r = subprocess.run([sys.executable, __file__, "child", n], env=env, capture_output=True, text=True, timeout=600)
        lines = [l for l in (r.stdout + r.stderr).splitlines() if "gemm" in l or "debug=" in l or "brute" in l]
        print("\n".join(lines[-6:]), flush=True)
