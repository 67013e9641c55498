"""GPU experiment: time the kNN search at config 2/3 sizes, look at the degree distribution of the 128-d graph and
time the Poisson kernels on it.  Not part of the product."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import graphlearning_b200 as gl
from graphlearning_b200 import knn_gpu, device as gdev
from oracle import gl_oracle as orc

def timed(f, *a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(*a, **k); torch.cuda.synchronize()
    return r, time.perf_counter() - t0

for (n, d, k) in ((70000, 128, 10), (60000, 512, 20)):
    X, labels = orc.synthetic_blobs(n, d, c=10, seed=0)
    X = X.astype(np.float64)
    knn_gpu.knnsearch_gpu(X[:4096], k + 1)                      # warm up (module load)
    (ind, dist), t = timed(knn_gpu.knnsearch_gpu, X, k + 1)
    print("knn n=%d d=%d k=%d: %.3f s end to end (host in/out), %d fallback rows, %.1f TFLOP/s on 2n^2d" % (
        n, d, k, t, knn_gpu.last_stats["fallback_rows"], 2.0 * n * n * d / t / 1e12), flush=True)
    W = orc.knn_weights(ind, dist, k)
    deg = np.diff(W.indptr)
    print("  graph: nnz=%d rows: min %d mean %.1f p99 %d max %d" % (W.nnz, deg.min(), deg.mean(), np.percentile(deg, 99), deg.max()), flush=True)
    ti = orc.one_per_class(labels, rate=1, seed=0)
    src = orc.poisson_source(n, ti, labels[ti])[0]
    for kind in ("dataflow", "barrier", "auto"):
        op = gdev.PoissonOperator(W, kind=kind)
        Db = op.source_to_Db(src)
        u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
        best = 1e9
        for _ in range(3):
            u0.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); op.iterate(Db, 500, u0, u1); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print("  poisson %-8s -> %-8s gate %d fill %.3f  %.3f us/iter" % (kind, op.kind(10), op.gate(10), op.fill(10), best * 1e3 / 500), flush=True)
