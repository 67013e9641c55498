"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on
seeded synthetic inputs.  Run by hand in the build container:

    python -m oracle.make_golden

TEST INFRASTRUCTURE ONLY.  The reference cannot travel to the GPU box, so its outputs are
committed as small fixtures; tests/test_oracle.py pins oracle/gl_oracle.py against them and the
-m gpu tests pin the CUDA path against both.

Reference entry points exercised (file:line in /root/reference):
  weightmatrix.knnsearch   graphlearning/weightmatrix.py:297  (method='kdtree' and 'brute')
  weightmatrix.knn         graphlearning/weightmatrix.py:68   (all kernels, symmetrize on/off)
  graph.degree_vector / laplacian        graphlearning/graph.py:108,469
  ssl.poisson (GD + CG) / ssl.laplace    graphlearning/ssl.py:513,1106   via .fit / .predict
  utils.conjgrad           graphlearning/utils.py:483  (2-D and 1-D right-hand sides)
  lp_iterate_main / lip_iterate_main     c_code/lp_iterate.cpp:35,129  (through oracle/_ref)
"""
import os

import numpy as np
from scipy import sparse
from sklearn import datasets as skdata

from . import c_oracle
from . import gl_oracle as orc
from ._refimport import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def csr_fields(prefix, W):
    W = sparse.csr_matrix(W)
    W.sort_indices()
    return {prefix + "_data": W.data.astype(np.float64), prefix + "_indices": W.indices.astype(np.int32),
            prefix + "_indptr": W.indptr.astype(np.int32), prefix + "_shape": np.array(W.shape)}


def fit(model, train_ind, train_labels):
    u = model.fit(train_ind, train_labels)
    return np.array(u), np.array(model.predict())


def main():
    gl = load_reference()
    os.makedirs(OUT, exist_ok=True)

    # ---- config 1: two moons, 500 points -------------------------------------------
    X, labels = skdata.make_moons(n_samples=500, noise=0.1, random_state=0)
    ind, dist = gl.weightmatrix.knnsearch(X, 11, method="kdtree")
    W = gl.weightmatrix.knn(X, 10)
    Wd = gl.weightmatrix.knn(X, 10, symmetrize=False)
    train_ind = orc.one_per_class(labels, rate=5, seed=1)
    tl = labels[train_ind]
    out = dict(X=X, labels=labels, knn_ind=ind, knn_dist=dist, train_ind=train_ind)
    out.update(csr_fields("W", W)); out.update(csr_fields("Wd", Wd))
    out["u_gd"], out["p_gd"] = fit(gl.ssl.poisson(W, solver="gradient_descent"), train_ind, tl)
    out["u_gd_T80"], _ = fit(gl.ssl.poisson(W, solver="gradient_descent", min_iter=80, max_iter=80), train_ind, tl)
    out["u_gd_directed"], out["p_gd_directed"] = fit(gl.ssl.poisson(Wd, solver="gradient_descent"), train_ind, tl)
    out["u_cg"], out["p_cg"] = fit(gl.ssl.poisson(W), train_ind, tl)
    out["u_lap"], out["p_lap"] = fit(gl.ssl.laplace(W), train_ind, tl)
    out["u_lap_norm"], _ = fit(gl.ssl.laplace(W, normalization="normalized", tau=0.01), train_ind, tl)
    out["u_lap_ms"], _ = fit(gl.ssl.laplace(W, mean_shift=True), train_ind, tl)
    G = gl.graph.graph(W)
    out["deg"] = G.degree_vector()
    for nm in ("combinatorial", "randomwalk", "normalized"):
        out.update(csr_fields("L_" + nm, G.laplacian(normalization=nm)))
    np.savez_compressed(os.path.join(OUT, "twomoons500.npz"), **out)
    print("twomoons500", {k: np.shape(v) for k, v in out.items() if k.startswith("u_")})

    # ---- blobs 2000 x 16 (scaled-down config 2) --------------------------------------------------
    Xb, lb = orc.synthetic_blobs(2000, 16, c=10, seed=0)
    Xb64 = Xb.astype(np.float64)
    ind_kd, dist_kd = gl.weightmatrix.knnsearch(Xb64, 11, method="kdtree")
    ind_br, dist_br = gl.weightmatrix.knnsearch(Xb64, 11, method="brute")
    assert np.array_equal(ind_kd, ind_br)
    Wb = gl.weightmatrix.knn(Xb64, 10, knn_data=(ind_kd, dist_kd))
    tb = orc.one_per_class(lb, rate=1, seed=0)
    out = dict(X=Xb, labels=lb, knn_ind=ind_kd.astype(np.int32), knn_dist=dist_kd, train_ind=tb)
    out.update(csr_fields("W", Wb))
    for T in (50, 200):
        out["u_gd_T%d" % T], out["p_gd_T%d" % T] = fit(
            gl.ssl.poisson(Wb, solver="gradient_descent", min_iter=T, max_iter=T), tb, lb[tb])
    out["u_cg"], out["p_cg"] = fit(gl.ssl.poisson(Wb), tb, lb[tb])
    t5 = orc.one_per_class(lb, rate=5, seed=0)
    out["train_ind5"] = t5
    out["u_lap"], out["p_lap"] = fit(gl.ssl.laplace(Wb), t5, lb[t5])
    # angular similarity search (weightmatrix.py:344-345)
    ia, da = gl.weightmatrix.knnsearch(Xb64, 11, method="kdtree", similarity="angular")
    out["knn_ind_angular"] = ia.astype(np.int32); out["knn_dist_angular"] = da
    np.savez_compressed(os.path.join(OUT, "blobs2000.npz"), **out)
    print("blobs2000 nnz", Wb.nnz)

    # ---- every kernel / symmetrisation rule of weightmatrix.knn on 300 points --------------------
    Xs, ls = orc.synthetic_blobs(300, 8, c=3, seed=3)
    Xs64 = Xs.astype(np.float64)
    inds, dists = gl.weightmatrix.knnsearch(Xs64, 8, method="kdtree")
    out = dict(X=Xs, labels=ls, knn_ind=inds.astype(np.int32), knn_dist=dists)
    for kernel in ("gaussian", "uniform", "symgaussian", "distance", "singular"):
        for sym in (True, False):
            Wk = gl.weightmatrix.knn(Xs64, 7, kernel=kernel, symmetrize=sym, knn_data=(inds.copy(), dists.copy()))
            out.update(csr_fields("W_%s_%d" % (kernel, int(sym)), Wk))
    # conjgrad on SPD systems, 2-D and 1-D right-hand sides (examples/regression.py:32 uses 1-D)
    Wg = gl.weightmatrix.knn(Xs64, 7, knn_data=(inds.copy(), dists.copy()))
    A = (gl.graph.graph(Wg).laplacian() + 0.1 * sparse.identity(300)).tocsr()
    rng = np.random.default_rng(7)
    B = rng.normal(size=(300, 3)); b1 = rng.normal(size=300)
    out.update(csr_fields("A", A)); out["B"] = B; out["b1"] = b1
    out["cg_x"] = gl.utils.conjgrad(A, B, tol=1e-8)
    out["cg_x1"] = gl.utils.conjgrad(A, b1, tol=1e-8)
    out["cg_x_it5"] = gl.utils.conjgrad(A, B, max_iter=5, tol=1e-30)

    # ---- p-Laplace iterates through the compiled reference (oracle/_ref) -------------------------
    I, J, V = orc.ccode_triplets(Wg)          # graph.py:69-84 convention: I=row, J=col
    bdy = np.array([0, 50, 100, 150, 200, 250], dtype=np.int32)
    g = (ls[bdy] == 0).astype(np.float64)
    uu0 = np.full(300, g.max()); ul0 = np.full(300, g.min())
    uu0[bdy] = g; ul0[bdy] = g
    # reference call order: (uu, ul, II=self.J (neighbour), J=self.I (row), W)  graph.py:1276
    for T in (7, 8, 200):
        a, b = c_oracle.lp_iterate(uu0, ul0, J, I, V, bdy, g, 3.0, T, 1e-6, use_ref=True)
        out["lp_uu_T%d" % T] = a; out["lp_ul_T%d" % T] = b
    out["lip_u_T30"] = c_oracle.lip_iterate(np.zeros(300), J, I, V, bdy, g, 30, 1e-9, 0.5, 0.5, use_ref=True)
    out["lip_u_conv"] = c_oracle.lip_iterate(np.zeros(300), J, I, V, bdy, g, 100000, 1e-6, 0.5, 0.5, use_ref=True)
    out["bdy"] = bdy; out["g"] = g; out["cI"] = I; out["cJ"] = J; out["cV"] = V
    np.savez_compressed(os.path.join(OUT, "small300.npz"), **out)
    print("small300 done")


if __name__ == "__main__":
    main()
