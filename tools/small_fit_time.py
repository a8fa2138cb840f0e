import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
from bench import make_blobs
import annchor_b200 as ab
from annchor_b200.annchor import Annchor
X = make_blobs(6000,128,100,42)
ctx = ab.default_context()
for it in range(4):
    a = Annchor(X,"euclidean",ctx=ctx,n_anchors=30,n_neighbors=15,n_samples=5000,p_work=0.001)
    ctx.timer_start(); a.fit(); ms = ctx.timer_stop()
    print(it, round(ms,1), {k:round(v*1e3,1) for k,v in a.stage_times.items()}, flush=True)
    a._index.close()
