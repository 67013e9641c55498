"""GPU experiment: exact kNN search at the config-2 / config-3 sizes: fused tcgen05 + TMA search (default, n >= 16384 and
d >= 64), the unfused tensor-core path (GLB_KNN_FUSED=0) and the fp32 SIMT path (GLB_KNN_TC=0).  Prints host-to-host and
device time, fallback rows and whether the indices agree with an fp64 brute-force check on a row sample.  With `full`
the whole 70 000 x 128 result is compared with scipy's cKDTree(workers=-1) (the reference's exact branch,
weightmatrix.py:349-352).  Needs the -DGLB_EXPERIMENT library.  Not part of the product."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GLB200_LIB", os.path.join(ROOT, "graphlearning_b200", "lib", "libglb200_exp.so"))
import numpy as np, torch
from graphlearning_b200 import knn_gpu, _lib, device as gdev
from oracle import gl_oracle as orc


def device_ms(X, k):
    """CUDA-event time of glb_knn_search on device-resident features"""
    Xd = torch.from_numpy(X).cuda()
    n = X.shape[0]
    ind = torch.empty((n, k), dtype=torch.int64, device="cuda"); dist = torch.empty((n, k), dtype=torch.float64, device="cuda")
    best = 1e9
    for _ in range(3):
        nl, nf = ctypes.c_int(0), ctypes.c_int(0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.call("glb_knn_search", gdev.ptr(Xd), n, X.shape[1], k, gdev.ptr(ind), gdev.ptr(dist), ctypes.byref(nl), ctypes.byref(nf), gdev.cur_stream())
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, ind.cpu().numpy(), nf.value, nl.value


configs = ((70000, 128, 11), (60000, 512, 21), (20000, 64, 11), (30000, 200, 11))
modes = (("fused", {}), ("unfused tc", {"GLB_KNN_FUSED": "0"}), ("fp32 simt", {"GLB_KNN_TC": "0"}))
for n, d, k in configs:
    X, _ = orc.synthetic_blobs(n, d, c=10, seed=0)
    X = X.astype(np.float64)
    rows = np.random.default_rng(1).choice(n, 200, replace=False)
    refs = []
    for r in rows:
        d2 = ((X - X[r]) ** 2).sum(1)
        refs.append(np.lexsort((np.arange(n), d2))[:k])
    refs = np.array(refs)
    first = None
    for name, env in modes:
        for key in ("GLB_KNN_FUSED", "GLB_KNN_TC"):
            os.environ.pop(key, None)
        os.environ.update(env)
        try:
            knn_gpu.knnsearch_gpu(X[:4096], k)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ind, dist = knn_gpu.knnsearch_gpu(X, k)
            torch.cuda.synchronize(); t_host = time.perf_counter() - t0
            ms, ind_d, nf, nl = device_ms(X, k)
            ok = bool(np.array_equal(refs, ind[rows])) and bool(np.array_equal(ind, ind_d))
            if first is None:
                first = ind
            print("%-10s n=%d d=%d k=%d: host to host %.4f s, device %.3f ms (%.1f TFLOP/s on 2n^2d), fallback rows %d, launches %d, "
                  "exact on 200 sampled rows: %s, identical to first mode: %s" % (name, n, d, k, t_host, ms, 2.0 * n * n * d / (ms * 1e-3) / 1e12, nf, nl, ok,
                                                                                bool(np.array_equal(ind, first))), flush=True)
        except Exception as e:
            print("%-10s n=%d d=%d k=%d FAILED: %r" % (name, n, d, k, e), flush=True)
    if "full" in sys.argv and (n, d) == (70000, 128):
        from scipy import spatial
        t0 = time.perf_counter()
        ref_dist, ref_ind = spatial.cKDTree(X).query(X, k=k, workers=-1)
        print("full-matrix parity n=%d d=%d: cKDTree(workers=-1) %.1f s on %d cores; indices identical: %s; max |dist - ref| = %.3e" % (
            n, d, time.perf_counter() - t0, os.cpu_count(), bool(np.array_equal(ref_ind, first)), float(np.abs(ref_dist - dist).max())), flush=True)
