"""Parity of the CUDA kNN search (knn.cu through the C-ABI) with the reference's exact searches.

Bar (north_star): bit-exact neighbour indices against knnsearch(method='kdtree'|'brute') on duplicate-free input;
distances are fp64 and may differ from scipy's in the last bits (different summation order): rtol 1e-12."""
import numpy as np
import pytest

from oracle import gl_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gl():
    import graphlearning_b200 as gl
    return gl


@pytest.fixture(scope="module")
def kg():
    from graphlearning_b200 import knn_gpu
    return knn_gpu


def check(ind, dist, ref_ind, ref_dist):
    assert ind.dtype == np.int64 and dist.dtype == np.float64
    assert np.array_equal(ind, ref_ind)
    assert np.allclose(dist, ref_dist, rtol=1e-12, atol=1e-300)


def test_goldens_from_the_reference(kg, moons, blobs, small):
    for g in (moons, blobs, small):
        k = g["knn_ind"].shape[1]
        ind, dist = kg.knnsearch_gpu(g["X"], k)
        check(ind, dist, g["knn_ind"], g["knn_dist"])
        assert kg.last_stats["launches"] > 0
    ind, dist = kg.knnsearch_gpu(blobs["X"], blobs["knn_ind_angular"].shape[1], similarity="angular")
    check(ind, dist, blobs["knn_ind_angular"], blobs["knn_dist_angular"])


def test_knnsearch_dispatch(gl, blobs, moons):
    """weightmatrix.knnsearch: d <= 5 -> cKDTree as the reference, otherwise (and for 'brute'/'annoy') the GPU."""
    k = blobs["knn_ind"].shape[1]
    for method in (None, "brute", "annoy"):
        ind, dist = gl.weightmatrix.knnsearch(blobs["X"], k, method=method)
        check(ind, dist, blobs["knn_ind"], blobs["knn_dist"])
    ind, _ = gl.weightmatrix.knnsearch(moons["X"], 11, method="brute")
    assert np.array_equal(ind, moons["knn_ind"])
    W = gl.weightmatrix.knn(blobs["X"], k - 1)
    Wref = blobs.csr("W")
    assert np.array_equal(W.indptr, Wref.indptr) and np.array_equal(W.indices, Wref.indices)
    assert np.allclose(W.data, Wref.data, rtol=1e-11)


@pytest.mark.parametrize("n,d,k", [(5000, 128, 11), (3000, 17, 31), (1500, 300, 101), (700, 1, 5), (40, 6, 40)])
def test_shapes_against_fp64_brute_force(kg, n, d, k):
    X, _ = orc.synthetic_blobs(n, d, c=5, seed=n + d)
    X = X.astype(np.float64)
    ind, dist = kg.knnsearch_gpu(X, k)
    rows = np.random.default_rng(0).choice(n, min(n, 300), replace=False)
    ref_ind, ref_dist = orc.knnsearch_rows(X, rows, k)
    check(ind[rows], dist[rows], ref_ind, ref_dist)
    assert np.array_equal(ind[:, 0], np.arange(n)) and np.all(dist[:, 0] == 0)
    assert np.all(np.diff(dist, axis=1) >= 0)


def test_duplicates_and_degenerate_data_use_the_exact_fallback(kg):
    rng = np.random.default_rng(3)
    X = rng.normal(size=(600, 20))
    X[100:140] = X[5]                               # 41 copies of one point: equal distances, certificate fails there
    ind, dist = kg.knnsearch_gpu(X, 11)
    assert kg.last_stats["fallback_rows"] >= 40
    ref_ind, ref_dist = orc.knnsearch_rows(X, np.arange(600), 11)
    assert np.allclose(dist, ref_dist, rtol=1e-12, atol=0)      # tie ORDER among equal distances is unspecified in the reference
    same = np.all(np.diff(ref_dist, axis=1) > 0, axis=1)        # rows without ties: indices exact
    assert np.array_equal(ind[same], ref_ind[same])
    assert np.all(dist[100:140, :11] == 0) and np.all(np.sort(ind[100:140], axis=1)[:, 0] == 5)
    Xc = np.ones((300, 8))                           # all points equal
    ind, dist = kg.knnsearch_gpu(Xc, 4)
    assert not dist.any() and np.array_equal(ind, np.tile(np.arange(4), (300, 1)))


def test_bad_arguments(kg):
    from graphlearning_b200._lib import GlbError
    with pytest.raises(GlbError):
        kg.knnsearch_gpu(np.zeros((10, 3)), 11)
    with pytest.raises(GlbError, match="not supported"):
        kg.knnsearch_gpu(np.zeros((500, 3)), 200)


def test_full_size_config2(kg):
    """70k x 128, k+1 = 11 (BASELINE config 2): exact on a 300-row sample, structural properties on all rows."""
    X, _ = orc.synthetic_blobs(70000, 128, c=10, seed=0)
    X = X.astype(np.float64)
    ind, dist = kg.knnsearch_gpu(X, 11)
    assert kg.last_stats["fallback_rows"] < 70                 # the certificate holds for (nearly) every row
    rows = np.random.default_rng(1).choice(70000, 300, replace=False)
    ref_ind, ref_dist = orc.knnsearch_rows(X, rows, 11)
    check(ind[rows], dist[rows], ref_ind, ref_dist)
    assert np.array_equal(ind[:, 0], np.arange(70000))
    assert np.all(np.diff(dist, axis=1) > 0)
    # symmetry of the distance: d(i, j) reported by row i equals |x_i - x_j| recomputed
    i = rows[:50]
    rec = np.sqrt(((X[i][:, None, :] - X[ind[i]]) ** 2).sum(-1))
    assert np.allclose(rec, dist[i], rtol=1e-12)


@pytest.mark.parametrize("n,d,k,nrows", [(60000, 512, 21, 2000), (20000, 64, 11, 600), (17001, 100, 16, 600), (33000, 300, 51, 400)])
def test_fused_tensor_core_search(kg, n, d, k, nrows):
    """n >= 16384 and d >= 64: the fused TMA + tcgen05 search (thresholds from a sample, one filtered pass, the distance
    block never leaves the SM).  Config 3 (60 000 x 512, k + 1 = 21) on 2 000 rows; sizes that are no multiple of the
    128-point / 64-feature tiles; a k that needs the 128-candidate list."""
    X, _ = orc.synthetic_blobs(n, d, c=10, seed=n % 97)
    X = X.astype(np.float64)
    ind, dist = kg.knnsearch_gpu(X, k)
    assert kg.last_stats["fallback_rows"] < n // 500              # the certificate holds for (nearly) every row
    rows = np.random.default_rng(2).choice(n, nrows, replace=False)
    rows[:3] = (0, n - 1, n // 2)
    Xr = X[rows]
    # fp64 brute force for the sampled rows (|x|^2 + |y|^2 - 2xy picks a shortlist, exact differences rank it)
    d2 = (Xr ** 2).sum(1)[:, None] + (X ** 2).sum(1)[None, :] - 2.0 * Xr @ X.T
    short = np.argpartition(d2, 4 * k, axis=1)[:, :4 * k]
    for t, i in enumerate(rows):
        ex = np.linalg.norm(X[short[t]] - X[i], axis=1)
        o = np.lexsort((short[t], ex))[:k]
        assert np.array_equal(ind[i], short[t][o]), (i, ind[i], short[t][o])
        assert np.allclose(dist[i], ex[o], rtol=1e-12, atol=1e-300)
    assert np.array_equal(ind[:, 0], np.arange(n)) and np.all(dist[:, 0] == 0)
    assert np.all(np.diff(dist, axis=1) >= 0)


def test_weight_matrix_assembly_on_device_is_bit_identical(gl, moons, blobs, small, monkeypatch):
    """weightmatrix.knn: COO -> CSR, symmetrisation by the kernel's rule, zero diagonal on the device (knn_graph.cu)
    against the goldens of the reference and against the scipy expressions (weightmatrix.py:166-186), bit for bit."""
    from scipy import sparse
    wm = gl.weightmatrix

    def same(A, B):
        A = sparse.csr_matrix(A); B = sparse.csr_matrix(B); B.sort_indices()
        assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices) and np.array_equal(A.data, B.data)

    monkeypatch.setattr(wm, "_device_assembly_min_n", 0)
    same(wm.knn(None, 10, knn_data=(moons["knn_ind"], moons["knn_dist"])), moons.csr("W"))
    same(wm.knn(None, 10, symmetrize=False, knn_data=(moons["knn_ind"], moons["knn_dist"])), moons.csr("Wd"))
    same(wm.knn(None, 7, knn_data=(small["knn_ind"].astype(np.int64), small["knn_dist"])), small.csr("W_gaussian_1"))
    same(wm.knn(None, 7, symmetrize=False, knn_data=(small["knn_ind"].astype(np.int64), small["knn_dist"])), small.csr("W_gaussian_0"))
    k = blobs["knn_ind"].shape[1]
    same(wm.knn(None, k - 1, knn_data=(blobs["knn_ind"], blobs["knn_dist"])), blobs.csr("W"))
    # a full-size graph: device path vs the scipy path of the same function (the reference's `eta` branch divides an
    # (n,k) array by an (n,) one, weightmatrix.py:162-164, and fails in numpy broadcasting - mirrored, not exercised)
    X, _ = orc.synthetic_blobs(70000, 8, c=10, seed=0)
    ind, dist = orc.knnsearch(X.astype(np.float64), 11, method="kdtree")
    for kw in ({}, {"symmetrize": False}):
        Wd = wm.knn(None, 10, knn_data=(ind, dist), **kw)
        monkeypatch.setattr(wm, "_device_assembly_min_n", 10 ** 9)
        Wh = wm.knn(None, 10, knn_data=(ind, dist), **kw)
        monkeypatch.setattr(wm, "_device_assembly_min_n", 0)
        same(Wd, Wh)
    assert (Wd != Wd.T).nnz > 0                                      # the last one is the directed graph
    # every other kernel: sparse_max for distance / uniform / singular, the symgaussian rule (weightmatrix.py:176-181),
    # against the goldens of the reference and against the scipy path at full size; a user kernel `eta` keeps the rule of `kernel`
    for kernel in ("uniform", "symgaussian", "distance", "singular"):
        for sym in (True, False):
            same(wm.knn(None, 7, kernel=kernel, symmetrize=sym, knn_data=(small["knn_ind"].astype(np.int64), small["knn_dist"])),
                 small.csr("W_%s_%d" % (kernel, int(sym))))
        Wd = wm.knn(None, 10, kernel=kernel, knn_data=(ind, dist))
        monkeypatch.setattr(wm, "_device_assembly_min_n", 10 ** 9)
        Wh = wm.knn(None, 10, kernel=kernel, knn_data=(ind, dist))
        monkeypatch.setattr(wm, "_device_assembly_min_n", 0)
        same(Wd, Wh)
    # a repeated column inside a row (the COO conversion sums duplicates first), one-directional edges, an exact zero weight
    dup_ind = np.array([[0, 1, 1], [1, 2, 0], [2, 0, 0], [3, 3, 1]])
    dup_w = np.array([[0.0, 0.25, 0.5], [0.0, 0.3, 0.9], [0.0, 0.7, 0.0], [0.0, 0.0, 0.4]])
    for rule, kernel in ((1, "gaussian"), (2, "uniform"), (3, "symgaussian")):
        got = wm._assemble_on_device(dup_ind, dup_w, 4, 3, rule)
        W0 = sparse.coo_matrix((dup_w.flatten(), (np.repeat(np.arange(4), 3), dup_ind.flatten())), shape=(4, 4)).tocsr()
        if rule == 1:
            want = (W0 + W0.T) / 2
        elif rule == 2:
            want = wm.sparse_max(W0, W0.transpose())
        else:
            want = W0 + W0.T.multiply(W0.T > W0) - W0.multiply(W0.T > W0)
        want = sparse.csr_matrix(want); want.setdiag(0); want.eliminate_zeros()
        same(got, want)
    with pytest.raises(Exception):
        wm.knn(None, 3, knn_data=(np.array([[0, 5, 1], [1, 0, 2], [2, 1, 0]]), np.ones((3, 3))))   # index out of range
