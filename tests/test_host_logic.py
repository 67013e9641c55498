"""Host-side mirror of the reference interface (no GPU): weight matrices, graph, predict, accuracy, trainsets."""
import numpy as np
import pytest
from scipy import sparse

import graphlearning_b200 as gl
from oracle import gl_oracle as orc


def same_csr(A, B):
    A = sparse.csr_matrix(A); B = sparse.csr_matrix(B)
    A.sort_indices(); B.sort_indices()
    assert np.array_equal(A.indptr, B.indptr) and np.array_equal(A.indices, B.indices)
    assert np.array_equal(A.data, B.data)


@pytest.mark.parametrize("kernel", ["gaussian", "uniform", "symgaussian", "distance", "singular"])
@pytest.mark.parametrize("sym", [True, False])
def test_knn_weights_match_reference(small, kernel, sym):
    W = gl.weightmatrix.knn(None, 7, kernel=kernel, symmetrize=sym,
                            knn_data=(small["knn_ind"].copy(), small["knn_dist"].copy()))
    same_csr(W, small.csr("W_%s_%d" % (kernel, int(sym))))


def test_knn_kdtree_path_matches_reference(moons):
    ind, dist = gl.weightmatrix.knnsearch(moons["X"], 11)          # d=2 -> kdtree, as in the reference
    assert np.array_equal(ind, moons["knn_ind"]) and np.array_equal(dist, moons["knn_dist"])
    same_csr(gl.weightmatrix.knn(moons["X"], 10), moons.csr("W"))
    same_csr(gl.weightmatrix.knn(moons["X"], 10, symmetrize=False), moons.csr("Wd"))


def test_graph_degree(moons):
    """degree vector on the host as the reference; the Laplacians are assembled on the device (tests/test_laplace_gpu.py)."""
    G = gl.graph(moons.csr("W"))
    assert G.num_nodes == 500
    assert np.array_equal(G.degree_vector(), moons["deg"])
    with pytest.raises(ValueError):
        G.laplacian(normalization="bogus")


def test_predict_and_accuracy(moons):
    m = gl.ssl.poisson(moons.csr("W"), solver="gradient_descent")
    with pytest.raises(RuntimeError):
        m.predict()
    m.prob = moons["u_gd"]; m.fitted = True
    assert np.array_equal(m.predict(), moons["p_gd"])
    acc = gl.ssl.ssl_accuracy(m.predict(), moons["labels"], moons["train_ind"])
    assert acc == orc.ssl_accuracy(moons["p_gd"], moons["labels"], moons["train_ind"])
    assert acc > 95


def test_invalid_options():
    with pytest.raises(ValueError):
        gl.ssl.poisson(None, solver="nope")
    with pytest.raises(ValueError):
        gl.weightmatrix.knnsearch(np.zeros((4, 2)), 2, method="bogus")
    with pytest.raises(ValueError):
        gl.weightmatrix.knnsearch(np.zeros((4, 2)), 2, similarity="hamming")
    with pytest.raises(RuntimeError):
        gl.ssl.poisson(None, solver="gradient_descent").fit(np.array([0]), np.array([0]))


def test_trainsets_generate_is_stratified_and_seeded():
    labels = np.repeat(np.arange(4), 25)
    a = gl.trainsets.generate(labels, rate=3, seed=5)
    b = gl.trainsets.generate(labels, rate=3, seed=5)
    assert np.array_equal(a, b) and len(a) == 12
    assert np.array_equal(np.bincount(labels[a]), [3, 3, 3, 3])
    sets = gl.trainsets.generate(labels, rate=2, num_trials=3, seed=1)
    assert len(sets) == 3


def test_labels_to_onehot_widens():
    oh = gl.utils.labels_to_onehot(np.array([0, 3, 1]), 2)
    assert oh.shape == (3, 4) and oh.sum() == 3


def test_drop_in_package_name():
    """`import graphlearning as gl` (the reference's package name) exposes the hot-path API the examples use
    (SURVEY.md appendix C): weightmatrix.knn/knnsearch, graph, ssl.poisson/laplace/plaplace/amle/ssl_accuracy,
    trainsets.generate, utils.conjgrad/labels_to_onehot, clustering.spectral/clustering_accuracy."""
    import graphlearning as gl
    import graphlearning.ssl as gssl
    for mod, names in ((gl.weightmatrix, ("knn", "knnsearch")), (gl.ssl, ("poisson", "laplace", "plaplace", "amle", "ssl_accuracy")),
                       (gl.trainsets, ("generate",)), (gl.utils, ("conjgrad", "labels_to_onehot", "randomized_svd")),
                       (gl.clustering, ("spectral", "clustering_accuracy", "purity"))):
        for n in names:
            assert callable(getattr(mod, n)), n
    assert gssl is gl.ssl and callable(gl.graph)
    G = gl.graph(__import__("scipy.sparse", fromlist=["x"]).identity(4, format="csr"))
    for n in ("degree_vector", "degree_matrix", "laplacian", "eigen_decomp", "plaplace", "amle"):
        assert callable(getattr(G, n)), n


def test_knn_data_on_disk_format(tmp_path, monkeypatch):
    """knnsearch(dataset=...) stores J/D in ./knn_data/<dataset>_<metric>.npz and knn('<dataset>', ...) /
    load_knn_data read it back (reference weightmatrix.py:416-427, 431-467)."""
    import graphlearning_b200 as gl
    monkeypatch.setattr(gl.weightmatrix, "knn_dir", str(tmp_path / "knn_data"))
    X = np.random.default_rng(0).normal(size=(200, 3))           # d <= 5: cKDTree, no GPU needed
    ind, dist = gl.weightmatrix.knnsearch(X, 8, dataset="Toy", metric="raw")
    J, D = gl.weightmatrix.load_knn_data("toy")
    assert np.array_equal(J, ind) and np.array_equal(D, dist)
    W1 = gl.weightmatrix.knn("toy", 7)
    W2 = gl.weightmatrix.knn(X, 7, knn_data=(ind, dist))
    assert (W1 != W2).nnz == 0
    with pytest.raises(FileNotFoundError):
        gl.weightmatrix.load_knn_data("absent")


def test_reweight_host_methods_against_reference_goldens(moons):
    """graph.reweight 'wnll' and 'properly' are pure host arithmetic (reference graph.py:436-462): bit-comparable with
    the goldens of the reference on CPU.  ('poisson' solves a linear system on the GPU: tests/test_cg_gpu.py.)"""
    from conftest import Golden
    rw = Golden("reweight")
    G = gl.graph(moons.csr("W"))
    ti, X = moons["train_ind"], moons["X"]
    for tag, kw in (("wnll", {}), ("properly", {"X": X})):
        Wr = sparse.csr_matrix(G.reweight(ti, method=tag, **kw)); Wr.sort_indices()
        assert np.array_equal(Wr.indices, rw["W_%s_indices" % tag]) and np.array_equal(Wr.indptr, rw["W_%s_indptr" % tag])
        assert np.allclose(Wr.data, rw["W_%s_data" % tag], rtol=1e-14, atol=0)
    with pytest.raises(ValueError):
        G.reweight(ti, method="nope")
    with pytest.raises(ValueError):
        G.reweight(ti, method="properly")


def test_ccode_triplets_match_reference_expressions(moons, blobs):
    """graph.I/J/V: the row-sorted COO triplets of graph.__ccode_init__ (graph.py:69-84); the canonical-CSR shortcut must give
    the arrays of the literal expressions, also with explicit zeros stored and for a non-canonical matrix."""
    for W in (moons.csr("W"), moons.csr("Wd"), blobs.csr("W")):
        G = gl.graph(W)
        I, J, V = orc.ccode_triplets(W)
        assert np.array_equal(G.I, I) and np.array_equal(G.J, J) and np.array_equal(G.V, V)
        assert G.I.dtype == np.int32 and G.J.dtype == np.int32 and G.V.dtype == np.float64
    W = moons.csr("W").copy()
    W.data[::7] = 0.0                                            # explicit zeros are dropped by sparse.find
    G = gl.graph(W)
    I, J, V = orc.ccode_triplets(W)
    assert np.array_equal(G.I, I) and np.array_equal(G.J, J) and np.array_equal(G.V, V)
    Wn = sparse.csr_matrix((np.array([1.0, 2.0, 3.0, 4.0]), np.array([2, 0, 1, 0]), np.array([0, 2, 3, 4])), shape=(3, 3))
    assert not Wn.has_sorted_indices
    G = gl.graph(Wn)
    I, J, V = orc.ccode_triplets(Wn)
    assert np.array_equal(G.I, I) and np.array_equal(G.J, J) and np.array_equal(G.V, V)


def test_clustering_scores():
    pred = np.array([1, 1, 0, 0, 2, 2, 2])
    true = np.array([5, 5, 7, 7, 9, 9, 7])
    assert abs(gl.clustering.clustering_accuracy(pred, true) - 100 * 6 / 7) < 1e-12
    overall, per_cluster = gl.clustering.purity(pred, true)                 # the reference returns both (clustering.py:550)
    assert abs(overall - 100 * 6 / 7) < 1e-12 and len(per_cluster) == len(np.unique(pred)) and per_cluster.max() <= 1
    assert gl.clustering.clustering_accuracy(true, true) == 100.0


def test_randomwalk_host_composition(moons, blobs, monkeypatch):
    """ssl.randomwalk (reference ssl.py:1731-1793) = Laplacian + host composition + ONE call of utils.conjgrad.  On CPU the
    two device calls (graph.laplacian, utils.conjgrad) are replaced by the oracle's (they have their own parity tests on the
    GPU), which pins everything around them to the golden of the reference."""
    from conftest import Golden, rel_err
    rwk = Golden("randomwalk")

    def cpu_conjgrad(A, b, x0=None, max_iter=1e5, tol=1e-10, return_info=False):
        x, it = orc.conjgrad(A, b, x0=x0, max_iter=max_iter, tol=tol, return_iters=True)
        return (x, (it, 0.0, 0)) if return_info else x

    monkeypatch.setattr(gl.utils, "conjgrad", cpu_conjgrad)
    monkeypatch.setattr(gl.graph, "laplacian", lambda self, normalization="combinatorial", alpha=1: orc.laplacian(self.weight_matrix, normalization))
    for name, g, tkey in (("moons", moons, "train_ind"), ("blobs", blobs, "train_ind5")):
        ti, labels = g[tkey], g["labels"]
        m = gl.ssl.randomwalk(g.csr("W"))
        u = m.fit(ti, labels[ti])
        assert rel_err(u, rwk[name + "_u"]) < 1e-12 and np.array_equal(m.predict(), rwk[name + "_pred"])
        assert m.accuracy_filename == "_randomwalk" and m.iterations > 0


# ---- dataflow kernel slabs (host-side builder of csrc/poisson.cu, checked without a GPU) ---------------------------
def _slab_check(P, c, grid):
    import ctypes
    from graphlearning_b200 import _lib
    P = sparse.csr_matrix(P)
    rp = np.ascontiguousarray(P.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(P.indices, dtype=np.int32)
    v = np.ascontiguousarray(P.data, dtype=np.float32)
    out = np.zeros(4)
    _lib.call("glb_dataflow_slabs_check_host", ctypes.c_void_p(rp.ctypes.data), ctypes.c_void_p(ci.ctypes.data),
              ctypes.c_void_p(v.ctypes.data), P.shape[0], c, grid, ctypes.c_void_p(out.ctypes.data))
    return {"err": out[0], "steps": out[1], "fill": out[2], "bad_rows": out[3]}


@pytest.mark.parametrize("c,grid", [(10, 148), (1, 7), (3, 16), (17, 148), (40, 33)])
def test_dataflow_slabs_hold_every_entry_once(c, grid):
    """The sliced-ELL slabs of the dataflow kernel, walked on the host exactly as the kernel walks them, give y = P x:
    every stored entry of P appears once, in the lane group of its row, padding entries carry the value 0, every row
    is stored by exactly one slot.  Includes empty rows and hub rows split over several warps."""
    rng = np.random.default_rng(c)
    n = 5000
    rows = np.repeat(np.arange(n), 9); cols = rng.integers(0, n, n * 9)
    hub = rng.integers(0, n, 3)                                     # three hubs: rows of 40, 300 and 2000 nonzeros
    rows = np.concatenate([rows] + [np.full(m, h) for h, m in zip(hub, (40, 300, 2000))])
    cols = np.concatenate([cols] + [rng.choice(n, m, replace=False) for m in (40, 300, 2000)])
    P = sparse.coo_matrix((rng.random(len(rows)), (rows, cols)), shape=(n, n)).tocsr()
    P = sparse.csr_matrix(P + P.T)
    keep = np.ones(n, bool); keep[rng.integers(0, n, 50)] = False   # 50 empty rows (and columns)
    D = sparse.diags(keep.astype(float)); P = sparse.csr_matrix(D @ P @ D); P.eliminate_zeros()
    r = _slab_check(P, c, grid)
    assert r["bad_rows"] == 0
    assert r["err"] < 1e-12
    assert 0.3 < r["fill"] <= 1.0


def test_pinned_pool_background_fill_and_recycling(monkeypatch):
    """device._PinnedPool without a GPU: the allocator and torch are stand-ins.  A miss returns an ordinary array at once and
    pins two buffers of that size on a background thread; later requests come from the pool; a buffer returns when the array
    and every view of it are gone; failures of the allocator never fail a request."""
    import ctypes
    import gc
    import types
    from graphlearning_b200 import device, _lib

    class FakeLib:
        def __init__(self):
            self.live, self.fail = {}, False

        def glb_host_alloc(self, nbytes, pp):
            if self.fail:
                return 2
            buf = ctypes.create_string_buffer(nbytes.value)
            addr = ctypes.addressof(buf)
            self.live[addr] = buf
            pp._obj.value = addr
            return 0

        def glb_host_free(self, p):
            self.live.pop(p.value, None)
            return 0

    fake = FakeLib()
    monkeypatch.setattr(_lib, "load", lambda: fake)
    cuda = types.SimpleNamespace(current_device=lambda: 0, set_device=lambda d: None)
    monkeypatch.setattr(device, "_torch", lambda: types.SimpleNamespace(cuda=cuda))
    pool = device._PinnedPool(keep=2, cap_bytes=10 << 20)
    shape = (20000, 10)                                         # 1.6 MB
    a = pool.empty(shape)
    assert a.shape == shape and a.flags.owndata                 # miss: pageable numpy memory, nothing waited for
    pool.wait()
    nbytes = a.nbytes
    assert len(pool.idle[nbytes]) == 2 and pool.total == 2 * nbytes
    b = pool.empty(shape); c = pool.empty(shape)
    assert not b.flags.owndata and not c.flags.owndata and len(pool.idle[nbytes]) == 0
    b[:] = 3.0
    view = b[5]
    del b
    gc.collect()
    assert len(pool.idle[nbytes]) == 0 and view[0] == 3.0       # a view keeps the buffer out of the pool
    del view, c
    gc.collect()
    assert len(pool.idle[nbytes]) == 2
    small = pool.empty((10, 10))
    assert small.flags.owndata and not any(pool.pending.values()) and 800 not in pool.idle      # below 64 KB nothing is pinned
    # keep = 2: a third returned buffer is freed, not kept
    d1, d2 = pool.empty(shape), pool.empty(shape)
    pool.empty(shape); pool.wait()                              # miss -> two more pinned
    d3, d4 = pool.empty(shape), pool.empty(shape)
    del d1, d2, d3, d4
    gc.collect()
    assert len(pool.idle[nbytes]) == 2 and pool.total == 2 * nbytes and len(fake.live) == 2
    # cap and allocator failures
    big = pool.empty((2000000, 10)); pool.wait()                # 160 MB > cap: never pinned
    assert big.flags.owndata and 160000000 not in pool.idle
    fake.fail = True
    e = pool.empty((30000, 10)); pool.wait()
    assert e.flags.owndata and pool.total == 2 * nbytes and not pool.idle.get(e.nbytes)


@pytest.mark.parametrize("c,n,sms", [(10, 6000, 148), (3, 6000, 148), (40, 4000, 148), (10, 60000, 4)])
def test_row_slab_stream_holds_every_entry_once(c, n, sms):
    """The entry stream of the row-slab kernel (csrc/slab.cu), walked on the host tile by tile, warp by warp as the kernel
    walks it, gives y = P x: every row is stored exactly once on its side of the boundary / interior split, padding lane
    groups gather the zero scratch row with the value 0, no warp part exceeds the warp's region of the stream buffers, and the
    folded slice order keeps the longest warp part near the mean.  Hub rows (a slice of their own), empty rows, halo columns;
    the last case has many tiles per CTA (4 SMs), which selects tiles of 32 slices."""
    import ctypes
    from graphlearning_b200 import _lib
    rng = np.random.default_rng(c + n)
    n_halo = 500
    rows = np.repeat(np.arange(n), 9); cols = rng.integers(0, n + n_halo, n * 9)
    hub = rng.choice(n, 2, replace=False)
    rows = np.concatenate([rows] + [np.full(k, h) for h, k in zip(hub, (150, 700))])
    cols = np.concatenate([cols] + [rng.choice(n + n_halo, k, replace=False) for k in (150, 700)])
    P = sparse.coo_matrix((rng.random(len(rows)).astype(np.float32), (rows, cols)), shape=(n, n + n_halo)).tocsr()
    P.sum_duplicates()
    keep = np.ones(n, bool); keep[rng.integers(0, n, 40)] = False
    P = sparse.csr_matrix(sparse.diags(keep.astype(np.float32)) @ P); P.eliminate_zeros()
    boundary = (P[:, n:].getnnz(axis=1) > 0).astype(np.uint8)                  # rows that read halo rows
    rp = np.ascontiguousarray(P.indptr, dtype=np.int32); ci = np.ascontiguousarray(P.indices, dtype=np.int32)
    va = np.ascontiguousarray(P.data, dtype=np.float32)
    out = np.zeros(8)
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    _lib.call("glb_slab_check_host", vp(rp), vp(ci), vp(va), n, n_halo, c, vp(boundary), sms, vp(out))
    err, bad, misplaced, fill, spc, region, longest_over_mean, tiles = out
    assert bad == 0 and misplaced == 0
    assert err < 1e-12
    assert 0.3 < fill <= 1.0
    assert spc == (32 if sms == 4 else 16)
    assert 2 * 8 * region + 1024 <= 227 * 1024                                  # two stream buffers of a CTA fit in shared memory
    if c == 10 and sms == 148:
        assert longest_over_mean < 8.0                                          # the hub's slice is the longest part
