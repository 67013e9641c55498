"""GPU experiment: schedules of the Gauss-Seidel sweep kernel (GLB_LIP_MODE bits: 1 level order, 2 producer poll, 4 lockstep, 8 sweeps overlap, 16 level counters)
on the 70k-node benchmark graph, same process.  Not part of the product."""
import ctypes, os, sys, time
import numpy as np
from scipy import sparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gl_oracle as orc
from graphlearning_b200 import _lib

n = 70000
X, labels = orc.synthetic_blobs(n, 8, c=10, seed=0)
ind, dist = orc.knnsearch(X.astype(np.float64), 11, method="kdtree")
W = sparse.csr_matrix(orc.knn_weights(ind, dist, 10))
I, J, V = orc.ccode_triplets(W)
ti = orc.one_per_class(labels, rate=5, seed=0).astype(np.int32)
val = (labels[ti] == 3).astype(np.float64)
p = lambda a: ctypes.c_void_p(a.ctypes.data)


def run(T, weighted):
    u = np.zeros(n); sw, nl = ctypes.c_int(), ctypes.c_int()
    t0 = time.perf_counter()
    _lib.call("glb_lip_iterate_host", p(u), p(J), p(I), p(V), p(ti), p(val), T, 1e-30, weighted, 0.5, 0.5, n, len(I), len(ti),
              ctypes.byref(sw), ctypes.byref(nl))
    return u, time.perf_counter() - t0


run(5, 0)
ref = {}
for weighted, T, modes in ((0, 1000, (0, 21, 17, 19, 0, 21)), (1, 200, (7, 21, 23, 7, 21))):
    for mode in modes:
        os.environ["GLB_LIP_MODE"] = str(mode)
        base = min(run(0, weighted)[1] for _ in range(3))
        res = [run(T, weighted) for _ in range(3)]
        t = min(r[1] for r in res)
        ref.setdefault(weighted, res[0][0])
        print("weighted=%d mode=%d  %.1f us/sweep  same bits as first=%s" % (weighted, mode, (t - base) / T * 1e6,
                                                                               bool(np.array_equal(res[0][0], ref[weighted]))), flush=True)
