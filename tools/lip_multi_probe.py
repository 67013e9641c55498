"""GPU experiment: one-vs-rest p-Laplace fit on the 70k-node graph (10 classes): batched sweep kernel vs ten single calls."""
import os, sys, time
import numpy as np
from scipy import sparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphlearning_b200 as gl
from oracle import gl_oracle as orc
n = 70000
X, labels = orc.synthetic_blobs(n, 8, c=10, seed=0)
ind, dist = orc.knnsearch(X.astype(np.float64), 11, method="kdtree")
W = sparse.csr_matrix(orc.knn_weights(ind, dist, 10))
ti = orc.one_per_class(labels, rate=5, seed=0)
G = gl.graph(W)
onehot = (labels[ti][:, None] == np.arange(10)[None, :]).astype(np.float64)
G.plaplace(ti, onehot[:, 0], 3, max_num_it=30)
for name, multi, single in (("plaplace p=3", lambda: G.plaplace(ti, onehot, 3), lambda k: G.plaplace(ti, onehot[:, k], 3)),
                            ("amle weighted tol 1e-3", lambda: G.amle(ti, onehot, tol=1e-3, max_num_it=300), lambda k: G.amle(ti, onehot[:, k], tol=1e-3, max_num_it=300))):
    t0 = time.perf_counter(); U = multi(); t1 = time.perf_counter(); sw = list(G.sweeps)
    cols = []; t2 = time.perf_counter()
    for k in range(10):
        cols.append(single(k))
    t3 = time.perf_counter()
    print("%s: batched %.3f s (sweeps %s), ten single calls %.3f s, identical %s" % (name, t1 - t0, sw, t3 - t2, bool(np.array_equal(U, np.stack(cols, 1)))), flush=True)
