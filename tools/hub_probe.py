"""GPU experiment: dataflow kernel statistics on the hub-heavy 128-d config-2 graph.  Not part of the product."""
import os, sys, time
os.environ["GLB_POISSON_STATS"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from graphlearning_b200 import knn_gpu, device as gdev
from oracle import gl_oracle as orc
n, d, k = 70000, 128, 10
X, labels = orc.synthetic_blobs(n, d, c=10, seed=0)
ind, dist = knn_gpu.knnsearch_gpu(X.astype(np.float64), k + 1)
W = orc.knn_weights(ind, dist, k)
ti = orc.one_per_class(labels, rate=1, seed=0)
src = orc.poisson_source(n, ti, labels[ti])[0]
for v in (sys.argv[1:] or ["512,16"]):
    for nopoll in ("0", "1"):
        os.environ["GLB_POISSON_VARIANT"] = v
        os.environ["GLB_POISSON_NOPOLL"] = nopoll
        op = gdev.PoissonOperator(W, kind="dataflow")
        Db = op.source_to_Db(src)
        u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
        for _ in range(2):
            u0.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); op.iterate(Db, 300, u0, u1); e1.record(); torch.cuda.synchronize()
        print("variant %s nopoll=%s: %.3f us/iter fill %.3f" % (v, nopoll, e0.elapsed_time(e1) * 1e3 / 300, op.fill(10)), flush=True)
