"""Parity of the CUDA p-Laplace / AMLE sweeps (plaplace.cu through the C-ABI) with the reference.

The three solvers of c_code/lp_iterate.cpp are sequential fp64 loops; the CUDA kernels keep the operation order
(per-row sums in stored order, no FMA contraction) and - for the Gauss-Seidel sweeps - the update order, so the
bar is BIT-EXACT equality with the goldens produced by the reference Python over the reference C code
(tests/golden/plaplace2000.npz, small300.npz) and with the plain-C oracle on larger seeded inputs."""
import ctypes

import numpy as np
import pytest
from scipy import sparse

from oracle import c_oracle
from oracle import gl_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gl():
    import graphlearning_b200 as gl
    return gl


def _ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def lip(u0, nbr, row, w, ind, val, T, tol, weighted, alpha=0.0, beta=1.0):
    from graphlearning_b200 import _lib
    u = np.ascontiguousarray(u0, dtype=np.float64).copy()
    nbr = np.ascontiguousarray(nbr, dtype=np.int32); row = np.ascontiguousarray(row, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    ind = np.ascontiguousarray(ind, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
    sw, nl = ctypes.c_int(-1), ctypes.c_int(0)
    _lib.call("glb_lip_iterate_host", _ptr(u), _ptr(nbr), _ptr(row), _ptr(w), _ptr(ind), _ptr(val), int(T), float(tol),
              int(weighted), float(alpha), float(beta), len(u), len(nbr), len(ind), ctypes.byref(sw), ctypes.byref(nl))
    assert nl.value > 0
    return u, sw.value


def lp(uu0, ul0, nbr, row, w, ind, val, p, T, tol):
    from graphlearning_b200 import _lib
    uu = np.ascontiguousarray(uu0, dtype=np.float64).copy(); ul = np.ascontiguousarray(ul0, dtype=np.float64).copy()
    nbr = np.ascontiguousarray(nbr, dtype=np.int32); row = np.ascontiguousarray(row, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    ind = np.ascontiguousarray(ind, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
    sw, nl = ctypes.c_int(-1), ctypes.c_int(0)
    _lib.call("glb_lp_iterate_host", _ptr(uu), _ptr(ul), _ptr(nbr), _ptr(row), _ptr(w), _ptr(ind), _ptr(val), float(p),
              int(T), float(tol), len(uu), len(nbr), len(ind), ctypes.byref(sw), ctypes.byref(nl))
    return uu, ul, sw.value


def test_small300_goldens_bit_exact(small):
    I, J, V, bdy, g = small["cI"], small["cJ"], small["cV"], small["bdy"], small["g"]
    uu0 = np.full(300, g.max()); ul0 = np.full(300, g.min()); uu0[bdy] = g; ul0[bdy] = g
    for T in (7, 8, 200):                                  # odd / even sweep counts: the pointer-swap parity of :116-123
        a, b, sw = lp(uu0, ul0, J, I, V, bdy, g, 3.0, T, 1e-6)
        assert np.array_equal(a, small["lp_uu_T%d" % T]) and np.array_equal(b, small["lp_ul_T%d" % T])
    u, sw = lip(np.zeros(300), J, I, V, bdy, g, 30, 1e-9, 0, 0.5, 0.5)
    assert np.array_equal(u, small["lip_u_T30"]) and sw == 30
    u, sw = lip(np.zeros(300), J, I, V, bdy, g, 100000, 1e-6, 0, 0.5, 0.5)
    assert np.array_equal(u, small["lip_u_conv"]) and sw < 100000


def test_plaplace2000_goldens_bit_exact(plap):
    I, J, V, ti, val = plap["cI"], plap["cJ"], plap["cV"], plap["train_ind"], plap["val"]
    n = 2000
    u, sw = lip(np.zeros(n), J, I, V, ti, val, 10 ** 6, 1e-6, 0, 0.5, 0.5)
    assert np.array_equal(u, plap["pl_fast_p3"])
    _, sw_ref = c_oracle.lip_iterate(np.zeros(n), J, I, V, ti.astype(np.int32), val, 10 ** 6, 1e-6, 0.5, 0.5)
    assert sw == sw_ref
    u, sw = lip(np.zeros(n), J, I, V, ti, val, 40, 1e-6, 0, 1 / 9, 1 - 1 / 9)
    assert np.array_equal(u, plap["pl_fast_p10_T40"]) and sw == 40
    uu = np.full(n, val.max()); ul = np.full(n, val.min()); uu[ti] = val; ul[ti] = val
    a, b, sw = lp(uu, ul, J, I, V, ti, val, 3.0, 10 ** 6, 1e-1)
    assert np.array_equal((a + b) / 2, plap["pl_slow_p3"])
    a, b, sw = lp(uu, ul, J, I, V, ti, val, 3.0, 101, 1e-9)
    assert np.array_equal((a + b) / 2, plap["pl_slow_p3_T101"]) and sw == 101
    u, sw = lip(np.zeros(n), J, I, V, ti, val, 1000, 1e-5, 1)
    assert np.array_equal(u, plap["amle_w"])
    u, sw = lip(np.zeros(n), J, I, V, ti, val, 25, 1e-5, 1)
    assert np.array_equal(u, plap["amle_w_T25"]) and sw == 25
    u, sw = lip(np.zeros(n), J, I, V, ti, val, 1000, 1e-5, 0, 0.0, 1.0)
    assert np.array_equal(u, plap["amle_u"])


def test_directed_graph_with_empty_rows(plap):
    """Empty rows: the reference reads u[I[start[i]]] - the next row's first neighbour - for min/max (:163, :223);
    the unweighted update then divides 0 by 0 (NaN), the weighted one copies that value."""
    ti, val = plap["train_ind"], plap["val"]
    u, _ = lip(np.zeros(2000), plap["dJ"], plap["dI"], plap["dV"], ti, val, 30, 1e-9, 1)
    assert np.array_equal(u, plap["amle_w_directed_T30"])
    u, _ = lip(np.zeros(2000), plap["dJ"], plap["dI"], plap["dV"], ti, val, 30, 1e-9, 0, 0.0, 1.0)
    assert np.array_equal(u, plap["amle_u_directed_T30"], equal_nan=True)


def test_python_api_matches_reference_python(gl, plap, blobs):
    W = blobs.csr("W")
    G = gl.graph(W)
    exact = np.array_equal(G.J, plap["cJ"]) and np.array_equal(G.I, plap["cI"])     # same stored order as the golden run
    same = (lambda a, b: np.array_equal(a, b)) if exact else (lambda a, b: np.allclose(a, b, rtol=1e-9, atol=1e-12))
    ti, val, labels = plap["train_ind"], plap["val"], blobs["labels"]
    assert same(G.plaplace(ti, val, 3), plap["pl_fast_p3"])
    assert G.sweeps > 0 and G.gpu_launches > 0
    assert same(G.plaplace(ti, val, 3, tol=1e-1, fast=False), plap["pl_slow_p3"])
    assert same(G.amle(ti, val, tol=1e-5, max_num_it=1000, weighted=True), plap["amle_w"])
    assert same(G.amle(ti, val == 1, tol=1e-5, max_num_it=1000, weighted=False), plap["amle_u"])   # bool values as ssl.fit passes
    m = gl.ssl.plaplace(W, p=3)
    u = m.fit(ti, labels[ti])
    assert same(u, plap["ssl_plaplace_p3"]) and np.array_equal(m.predict(), plap["ssl_plaplace_p3_pred"])
    m = gl.ssl.amle(W)
    u = m.fit(ti, labels[ti])
    assert same(u, plap["ssl_amle"]) and np.array_equal(m.predict(), plap["ssl_amle_pred"])
    assert gl.ssl.ssl_accuracy(m.predict(), labels, ti) > 90


def test_edge_cases(plap):
    I, J, V, ti, val = plap["cI"], plap["cJ"], plap["cV"], plap["train_ind"], plap["val"]
    n = 2000
    rng = np.random.default_rng(3)
    u0 = rng.normal(size=n)
    # T = 0: only the Dirichlet values are written
    u, sw = lip(u0, J, I, V, ti, val, 0, 1e-6, 0, 0.5, 0.5)
    ref = u0.copy(); ref[ti] = val
    assert np.array_equal(u, ref) and sw == 0
    # non-zero start, duplicated boundary index (the last value wins), one sweep
    ind2 = np.concatenate([ti, ti[:3]]); val2 = np.concatenate([val, [7.0, 8.0, 9.0]])
    for weighted in (0, 1):
        u, sw = lip(u0, J, I, V, ind2, val2, 3, 1e-6, weighted, 0.3, 0.7)
        if weighted:
            r, _ = c_oracle.lip_iterate_weighted(u0, J, I, V, ind2.astype(np.int32), val2, 3, 1e-6)
        else:
            r, _ = c_oracle.lip_iterate(u0, J, I, V, ind2.astype(np.int32), val2, 3, 1e-6, 0.3, 0.7)
        assert np.array_equal(u, r) and sw == 3
    # no boundary at all
    u, _ = lip(u0, J, I, V, np.zeros(0, np.int32), np.zeros(0), 5, 1e-6, 0, 0.5, 0.5)
    r, _ = c_oracle.lip_iterate(u0, J, I, V, np.zeros(0, np.int32), np.zeros(0), 5, 1e-6, 0.5, 0.5)
    assert np.array_equal(u, r)
    # rows not sorted -> argument error, no result
    from graphlearning_b200 import _lib
    with pytest.raises(_lib.GlbError):
        lip(u0, J, I[::-1].copy(), V, ti, val, 3, 1e-6, 0, 0.5, 0.5)
    with pytest.raises(_lib.GlbError):
        lip(u0, J + n, I, V, ti, val, 3, 1e-6, 0, 0.5, 0.5)


def test_rows_longer_than_the_local_cache():
    """A graph with hub rows (> 32 neighbours, beyond the per-lane cache of the bisection) and consecutive
    dependent rows inside one warp (a path graph: row i depends on row i-1 in the same sweep)."""
    n = 3000
    rng = np.random.default_rng(11)
    rows = [np.arange(n - 1), np.arange(1, n)]
    cols = [np.arange(1, n), np.arange(n - 1)]
    hub = rng.choice(n, 200, replace=False)
    for h in (17, 1234):
        rows += [np.full(200, h), hub]; cols += [hub, np.full(200, h)]
    r = np.concatenate(rows); c = np.concatenate(cols)
    keep = r != c
    W = sparse.csr_matrix((rng.uniform(0.1, 1.0, keep.sum()), (r[keep], c[keep])), shape=(n, n))
    W = sparse.csr_matrix((W + W.T) / 2)
    I, J, V = orc.ccode_triplets(W)
    ind = np.array([0, n // 2, n - 1], dtype=np.int32); val = np.array([0.0, 1.0, -1.0])
    for weighted in (0, 1):
        u, sw = lip(np.zeros(n), J, I, V, ind, val, 60, 1e-12, weighted, 0.4, 0.6)
        if weighted:
            ref, _ = c_oracle.lip_iterate_weighted(np.zeros(n), J, I, V, ind, val, 60, 1e-12)
        else:
            ref, _ = c_oracle.lip_iterate(np.zeros(n), J, I, V, ind, val, 60, 1e-12, 0.4, 0.6)
        assert np.array_equal(u, ref) and sw == 60


def test_70k_graph_against_oracle():
    """Full benchmark size (70k nodes, k=10): bit-exact against the plain-C oracle, and the size-independent
    properties of the solution - a discrete maximum principle (values inside the range of the boundary data) and
    idempotence (restarting from the converged iterate changes nothing beyond the tolerance)."""
    n = 70000
    X, labels = orc.synthetic_blobs(n, 8, c=10, seed=0)
    ind, dist = orc.knnsearch(X.astype(np.float64), 11, method="kdtree")
    W = sparse.csr_matrix(orc.knn_weights(ind, dist, 10))
    I, J, V = orc.ccode_triplets(W)
    ti = orc.one_per_class(labels, rate=5, seed=0).astype(np.int32)
    val = (labels[ti] == 3).astype(np.float64)
    u, sw = lip(np.zeros(n), J, I, V, ti, val, 40, 1e-5, 0, 0.5, 0.5)
    ref, sw_ref = c_oracle.lip_iterate(np.zeros(n), J, I, V, ti, val, 40, 1e-5, 0.5, 0.5)
    assert sw == sw_ref and np.array_equal(u, ref)
    uw, sw = lip(np.zeros(n), J, I, V, ti, val, 12, 1e-5, 1)
    refw, _ = c_oracle.lip_iterate_weighted(np.zeros(n), J, I, V, ti, val, 12, 1e-5)
    assert np.array_equal(uw, refw)
    assert u.min() >= 0.0 and u.max() <= 1.0 and uw.min() >= 0.0 and uw.max() <= 1.0
    uu = np.ones(n); ul = np.zeros(n); uu[ti] = val; ul[ti] = val
    a, b, sw = lp(uu, ul, J, I, V, ti, val, 3.0, 25, 1e-3)
    ra, rb, _ = c_oracle.lp_iterate(uu, ul, J, I, V, ti, val, 3.0, 25, 1e-3)
    assert np.array_equal(a, ra) and np.array_equal(b, rb)
    assert np.all(a >= b)                                   # upper barrier stays above the lower one
    conv, sw1 = lip(np.zeros(n), J, I, V, ti, val, 100000, 1e-4, 0, 0.5, 0.5)
    again, sw2 = lip(conv, J, I, V, ti, val, 100000, 1e-4, 0, 0.5, 0.5)
    assert sw2 == 22 and np.max(np.abs(again - conv)) < 22 * 1e-4      # `it > 20` forces 22 sweeps, each moving < tol


def test_batched_classes_equal_separate_calls(gl, plap, blobs):
    """glb_lip_iterate_multi_host: the c one-vs-rest right-hand sides in one launch; every column - values AND the sweep at
    which its own stopping rule fired - must equal the single-class call, and ssl.plaplace / ssl.amle (which use the
    batched path) must still match the reference goldens."""
    I, J, V, ti = plap["cI"], plap["cJ"], plap["cV"], plap["train_ind"]
    labels = blobs["labels"]
    onehot = (labels[ti][:, None] == np.unique(labels[ti])[None, :]).astype(np.float64)
    G = gl.graph(blobs.csr("W"))
    for weighted, tol, T in ((False, 1e-5, 1000), (True, 1e-3, 60)):
        U = G.amle(ti, onehot, tol=tol, max_num_it=T, weighted=weighted)
        sw_multi = list(G.sweeps)
        assert U.shape == (2000, onehot.shape[1]) and len(set(sw_multi)) > 1 or weighted     # classes stop at different sweeps
        for k in range(onehot.shape[1]):
            uk, sk = lip(np.zeros(2000), G.J, G.I, G.V, ti, onehot[:, k], T, tol, int(weighted), 0.0, 1.0)
            assert np.array_equal(U[:, k], uk) and sw_multi[k] == sk, (weighted, k, sw_multi[k], sk)
    m = gl.ssl.plaplace(blobs.csr("W"), p=3)
    u = m.fit(ti, labels[ti])
    exact = np.array_equal(G.J, plap["cJ"]) and np.array_equal(G.I, plap["cI"])
    assert np.array_equal(u, plap["ssl_plaplace_p3"]) if exact else np.allclose(u, plap["ssl_plaplace_p3"], rtol=1e-9, atol=1e-12)
    m = gl.ssl.amle(blobs.csr("W"))
    u = m.fit(ti, labels[ti])
    assert np.array_equal(u, plap["ssl_amle"]) if exact else np.allclose(u, plap["ssl_amle"], rtol=1e-9, atol=1e-12)
