"""Config 4 on the GPU: 70k-node k=10 graph, 50 eigenpairs of the normalised Laplacian (graph.eigen_decomp), plus the
raw SpMM / Gram / right-multiply kernel times at that block width.  Prints one JSON object."""
import json, os, sys, time
import numpy as np
from scipy import sparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import graphlearning_b200 as gl
from graphlearning_b200.spectral import BlockOps
from oracle import gl_oracle as orc

n = 70000
X, labels = orc.synthetic_blobs(n, 8, c=10, seed=0)
ind, dist = orc.knnsearch(X.astype(np.float64), 11, method="kdtree")
W = sparse.csr_matrix(orc.knn_weights(ind, dist, 10))
out = {"n": n, "nnz": int(W.nnz)}
G = gl.graph(W)
G.eigen_decomp(normalization="normalized", k=4)                         # warm-up
for method, kw in (("exact", {}), ("lowrank", {"q": 10})):
    G = gl.graph(W)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    vals, vecs = G.eigen_decomp(normalization="normalized", k=50, method=method, **kw)
    torch.cuda.synchronize(); t = time.perf_counter() - t0
    L = G.laplacian(normalization="normalized")
    out["eigen_decomp_k50_" + method] = {"seconds": t, "info": {k: (v if not isinstance(v, (np.floating, np.integer)) else v.item()) for k, v in G.eigen_info.items()},
                                         "max_residual_L": float(np.max(np.abs(L @ vecs - vecs * vals))), "vals_head": vals[:6].tolist(), "vals_tail": vals[-3:].tolist()}
deg = np.asarray(W.sum(axis=1)).ravel()
D = sparse.spdiags(deg ** -0.5, 0, n, n)
ops = BlockOps(sparse.csr_matrix(D @ W @ D))
for c in (10, 62, 100):
    Xd = ops.upload(np.random.default_rng(0).standard_normal((n, c))); Z = ops.new(c); Y = ops.new(c)
    S = np.random.default_rng(1).standard_normal((c, c))
    res = {}
    for name, fn in (("spmm", lambda: ops.spmm(Xd, c, out=Z)), ("spmm_fused", lambda: ops.spmm(Xd, c, out=Z, Y1=Y, beta=0.5, Y2=Z, gamma=0.1)),
                     ("gram", lambda: ops.gram(Xd, c, Xd, c)), ("right_mul", lambda: ops.right_mul(Xd, c, S, out=Z))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        res[name + "_us"] = e0.elapsed_time(e1) / 20 * 1e3
    ld = (c + 1) & ~1
    res["spmm_algorithmic_GBs"] = (W.nnz * 12 + (n + 1) * 4 + 2 * n * ld * 8) / (res["spmm_us"] * 1e-6) / 1e9
    res["gram_GFLOPs"] = 2.0 * n * c * c / (res["gram_us"] * 1e-6) / 1e9
    out["kernels_c%d" % c] = res
print(json.dumps(out))
