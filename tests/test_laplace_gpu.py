"""Device assembly of the graph Laplacians and of the Laplace-learning system (csrc/laplace.cu) against the reference's
scipy expressions (graphlearning/graph.py:469-513, ssl.py:1222-1255) and its goldens."""
import ctypes

import numpy as np
import pytest
from scipy import sparse

from oracle import gl_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gl():
    import graphlearning_b200 as g
    return g


def rel_err(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def scipy_laplacian(W, normalization):
    """the reference's expressions, literally (graph.py:493-503)"""
    n = W.shape[0]
    I = sparse.identity(n)
    d = W * np.ones(n)
    Dp = lambda p: sparse.spdiags(d ** p, 0, n, n).tocsr()
    if normalization == "combinatorial":
        return (Dp(1) - W).tocsr()
    if normalization == "randomwalk":
        return (I - Dp(-1) * W).tocsr()
    return (I - Dp(-0.5) * W * Dp(-0.5)).tocsr()


def assert_same_matrix(A, B, same_order):
    A = sparse.csr_matrix(A); B = sparse.csr_matrix(B)
    if same_order:                                   # stored order too: what a CSR product sums in
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)
    A = A.copy(); B = B.copy(); A.sort_indices(); B.sort_indices()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    assert np.array_equal(A.data, B.data)            # bit for bit


@pytest.mark.parametrize("nm", ["combinatorial", "randomwalk", "normalized"])
def test_laplacian_is_bit_identical_to_the_reference(gl, moons, blobs, nm):
    assert_same_matrix(gl.graph(moons.csr("W")).laplacian(normalization=nm), moons.csr("L_" + nm), same_order=False)
    for W in (moons.csr("W"), moons.csr("Wd"), blobs.csr("W")):          # symmetric, directed, 2000 nodes
        L = gl.graph(W).laplacian(normalization=nm)
        assert L.has_canonical_format
        # scipy leaves the random-walk Laplacian in an unsorted order; the other two come out canonical there too
        assert_same_matrix(L, scipy_laplacian(sparse.csr_matrix(W), nm), same_order=(nm != "randomwalk"))


def test_laplacian_with_self_loops_zero_results_and_70k_nodes(gl):
    from test_poisson_gpu import random_knn_graph
    rng = np.random.default_rng(0)
    W = random_knn_graph(3000, 6, seed=1).tolil()
    for i in rng.integers(0, 3000, 200):
        W[i, i] = rng.random()                        # stored diagonal entries
    W = sparse.csr_matrix(W)
    for nm in ("combinatorial", "randomwalk", "normalized"):
        assert_same_matrix(gl.graph(W).laplacian(normalization=nm), scipy_laplacian(W, nm), same_order=False)
    # a node whose only weight is its self loop: D - W has an exact zero there, which scipy does not store
    Z = sparse.csr_matrix(np.array([[2.0, 0, 0], [0, 0, 1.0], [0, 1.0, 0]]))
    L = gl.graph(Z).laplacian()
    assert_same_matrix(L, scipy_laplacian(Z, "combinatorial"), same_order=True)
    assert L.nnz == 4
    # coifmanlafon goes through the same kernel after the host's D W D
    Wm = random_knn_graph(500, 5, seed=2)
    d = Wm * np.ones(500)
    D = sparse.spdiags(d ** -1.0, 0, 500, 500).tocsr()
    assert_same_matrix(gl.graph(Wm).laplacian(normalization="coifmanlafon"), scipy_laplacian(sparse.csr_matrix(D * Wm * D), "randomwalk"),
                       same_order=False)
    Wb = random_knn_graph(70000, 10, seed=3)
    assert_same_matrix(gl.graph(Wb).laplacian(normalization="normalized"), scipy_laplacian(Wb, "normalized"), same_order=True)


def test_fused_laplace_fit_equals_host_assembly_plus_cg(gl, blobs):
    """ssl.laplace._fit_device (system assembled in HBM) against the reference's scipy assembly followed by the same CG,
    for every normalisation, a tau vector, labelled nodes given with negative indices; order = 2 and repeated labelled nodes
    take the host-assembly path."""
    W = blobs.csr("W")
    tb = blobs["train_ind5"]; tl = blobs["labels"][tb]
    n = W.shape[0]
    tau_vec = np.random.default_rng(1).random(n) * 0.01
    for kw in ({}, {"normalization": "randomwalk"}, {"normalization": "normalized", "tau": 0.02}, {"tau": tau_vec}):
        m = gl.ssl.laplace(W, **kw)
        u = m.fit(tb, tl)
        MAM, Mb, M, idx, F = m.system(tb, tl)
        if kw.get("normalization") == "randomwalk":
            # M A M of the random-walk Laplacian is not symmetric: CG does not converge on it - the reference runs into
            # max_iter = 1e5 with err ~ 0.06 and growing - and where 1e5 steps of an unstable recurrence end (max_iter, or a
            # NaN when some p.Ap cancels to zero) is rounding noise.  Parity on a bounded horizon, and no early stop.
            v_ref, it = orc.conjgrad(MAM, Mb, tol=1e-5, max_iter=3000, return_iters=True)
            v, (it_d, err_d, _) = gl.utils.conjgrad(MAM, Mb, tol=1e-5, max_iter=3000, return_info=True)
            assert it == it_d == 3000
            assert rel_err(v, v_ref) <= 1e-6
            assert m.iterations > 3000 and np.array_equal(u[tb], F)
            continue
        v, it = orc.conjgrad(MAM, Mb, tol=1e-5, return_iters=True)
        u_ref = np.zeros_like(u); u_ref[idx] = M * v; u_ref[tb] = F
        assert abs(m.iterations - it) <= 1 and m.gpu_launches > 0
        assert rel_err(u, u_ref) <= 1e-6
        assert np.array_equal(u[tb], F)
    u_neg = gl.ssl.laplace(W).fit(tb - n, tl)
    assert np.array_equal(u_neg, gl.ssl.laplace(W).fit(tb, tl))
    m2 = gl.ssl.laplace(W, order=2)
    u2 = m2.fit(tb, tl)
    s = m2.system(tb, tl)
    v2 = orc.conjgrad(s[0], s[1], tol=1e-5)
    assert rel_err(u2[s[3]], s[2] * v2) <= 1e-5
    dup = np.concatenate([tb, tb[:3]])
    u_dup = gl.ssl.laplace(W).fit(dup, blobs["labels"][dup])
    assert u_dup.shape == u.shape and np.isfinite(u_dup).all()


def test_laplace_fit_rejects_bad_input(gl, moons):
    from graphlearning_b200 import _lib
    W = moons.csr("W")
    rp = np.ascontiguousarray(W.indptr, np.int32); ci = np.ascontiguousarray(W.indices, np.int32); v = np.ascontiguousarray(W.data)
    d = W * np.ones(500)
    F = np.eye(2)
    u = np.empty((500, 2))
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    for ti in (np.array([0, 700], np.int64), np.array([-1, 3], np.int64)):
        with pytest.raises(_lib.GlbError):
            _lib.call("glb_laplace_fit_host", vp(rp), vp(ci), vp(v), 500, W.nnz, None, None, vp(d), None, vp(ti), 2, vp(F), 2, 1e-5,
                      vp(u), None, None, None, None)
