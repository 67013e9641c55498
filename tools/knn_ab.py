"""GPU experiment: exact kNN search at the config-2 / config-3 sizes with the distance blocks on the tensor cores
(GLB_KNN_TC=1, default for d >= 64) or on the fp32 SIMT kernel (GLB_KNN_TC=0).  Prints time, fallback rows and whether
the indices agree with an fp64 brute-force check on a row sample.  Not part of the product."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphlearning_b200 import knn_gpu
from oracle import gl_oracle as orc

for n, d, k in ((70000, 128, 11), (60000, 512, 21), (20000, 64, 11)):
    X, _ = orc.synthetic_blobs(n, d, c=10, seed=0)
    X = X.astype(np.float64)
    knn_gpu.knnsearch_gpu(X[:4096], k)
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ind, dist = knn_gpu.knnsearch_gpu(X, k)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    rows = np.random.default_rng(1).choice(n, 200, replace=False)
    ok = True
    for r in rows:
        d2 = ((X - X[r]) ** 2).sum(1)
        ref = np.lexsort((np.arange(n), d2))[:k]
        ok &= bool(np.array_equal(ref, ind[r]))
    print("TC=%s n=%d d=%d k=%d: %.4f s host to host, %.1f TFLOP/s on 2n^2d, fallback rows %s, launches %s, exact on 200 sampled rows: %s" % (
        os.environ.get("GLB_KNN_TC", "default"), n, d, k, best, 2.0 * n * n * d / best / 1e12, knn_gpu.last_stats.get("fallback_rows"),
        knn_gpu.last_stats.get("launches"), ok), flush=True)
