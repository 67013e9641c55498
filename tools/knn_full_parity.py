"""GPU check, run once per change of the search: the WHOLE index matrix of glb_knn_search at the BASELINE sizes against
an exact fp64 brute force on the device (torch.cdist without the matrix-multiply shortcut: sum((x_i - x_j)^2) in fp64, then
the k smallest by (distance, index)) - the ranking the reference's exact branches produce (weightmatrix.py:349-361).
With `kdtree` config 2 is also compared with scipy's cKDTree(workers=-1), the reference's own call (:351-352).
Test infrastructure, not part of the product.   usage: python tools/knn_full_parity.py [kdtree]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from graphlearning_b200 import knn_gpu
from oracle import gl_oracle as orc


def brute(X, k, block=2048):
    Xd = torch.from_numpy(X).cuda()
    n = X.shape[0]
    out = np.empty((n, k), dtype=np.int64)
    dist = np.empty((n, k), dtype=np.float64)
    for r0 in range(0, n, block):
        D = torch.cdist(Xd[r0:r0 + block], Xd, compute_mode="donot_use_mm_for_euclid_dist")
        d, i = torch.topk(D, k + 4, dim=1, largest=False, sorted=True)
        d = d.cpu().numpy(); i = i.cpu().numpy()
        order = np.lexsort((i, d), axis=1)[:, :k]                # ties (none expected) by index
        out[r0:r0 + block] = np.take_along_axis(i, order, 1)
        dist[r0:r0 + block] = np.take_along_axis(d, order, 1)
    return out, dist


for n, d, k in ((70000, 128, 11), (60000, 512, 21)):
    X, _ = orc.synthetic_blobs(n, d, c=10, seed=0)
    X = X.astype(np.float64)
    t0 = time.perf_counter(); ind, dist = knn_gpu.knnsearch_gpu(X, k); t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter(); ref, rdist = brute(X, k); t_ref = time.perf_counter() - t0
    bad = int((ind != ref).any(axis=1).sum())
    print("n=%d d=%d k=%d: search %.3f s (fallback rows %d), fp64 brute force %.1f s, rows with any differing index: %d of %d, "
          "max |dist - ref| = %.2e" % (n, d, k, t_gpu, knn_gpu.last_stats["fallback_rows"], t_ref, bad, n,
                                       float(np.abs(dist - rdist).max())), flush=True)
    if "kdtree" in sys.argv and d == 128:
        from scipy.spatial import cKDTree
        t0 = time.perf_counter()
        kd_dist, kd_ind = cKDTree(X).query(X, k=k, workers=-1)
        t_kd = time.perf_counter() - t0
        print("    scipy cKDTree(workers=-1, %d cores): %.1f s, rows with any differing index: %d of %d" % (
            os.cpu_count(), t_kd, int((ind != kd_ind).any(axis=1).sum()), n), flush=True)
