import sys, time, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from conftest import Golden
import graphlearning_b200 as gl
blobs = Golden("blobs2000")
W = blobs.csr("W"); tb = blobs["train_ind5"]; tl = blobs["labels"][tb]
m = gl.ssl.laplace(W, normalization="randomwalk")
MAM, Mb, M, idx, F = m.system(tb, tl)
for mi in (100, 1000, 3000, 10000, 44000, 44369, 100000):
    t0 = time.time()
    x, (it, err, nl) = gl.utils.conjgrad(MAM, Mb, max_iter=mi, tol=1e-5, return_info=True)
    print("max_iter %6d -> it %6d err %.6g  max|x| %.4g  finite %s  (%.1f s)" % (mi, it, err, np.abs(x).max(), np.isfinite(x).all(), time.time() - t0), flush=True)
