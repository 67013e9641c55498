"""Parity of the CUDA Poisson path (through the C-ABI) against the oracle and the reference-generated goldens.

Tolerance (north_star): label scores fp32 within 1e-5 relative = max|u - u_ref| / max|u_ref| <= 1e-5 after the
same iteration count; predicted labels identical; integer/index work (transpose pattern, T) bit-exact.
"""
import ctypes

import numpy as np
import pytest
from scipy import sparse

from conftest import rel_err
from oracle import c_oracle
from oracle import gl_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def gl():
    import graphlearning_b200 as gl
    return gl


@pytest.fixture(scope="module")
def dev():
    from graphlearning_b200 import device
    return device


def random_knn_graph(n, k, seed=0, symmetric=True):
    """kNN-shaped random graph without a search: k random neighbours per row, gaussian-like weights."""
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), k)
    cols = rng.integers(0, n, n * k)
    w = np.exp(-4 * rng.random(n * k))
    W = sparse.coo_matrix((w, (rows, cols)), shape=(n, n)).tocsr()
    if symmetric:
        W = (W + W.T) / 2
    W = sparse.csr_matrix(W)
    W.setdiag(0)
    W.eliminate_zeros()
    return W


# ---- through the reference-facing API (host buffers in, host buffers out) ----------------------------
def test_two_moons_default_stopping_rule(gl, moons):
    ti = moons["train_ind"]; tl = moons["labels"][ti]
    _, T_ref = orc.poisson_gd(moons.csr("W"), ti, tl, return_iters=True)
    m = gl.ssl.poisson(moons.csr("W"), solver="gradient_descent")
    u = m.fit(ti, tl)
    assert m.iterations == T_ref                       # integer work: exact
    assert u.shape == (500, 2) and u.dtype == np.float64
    assert rel_err(u, moons["u_gd"]) <= TOL
    assert np.array_equal(m.predict(), moons["p_gd"])
    assert m.gpu_launches > 0


def test_all_labels_prints_the_accuracy_of_every_iteration(gl, moons, capsys):
    """fit(..., all_labels=labels): one line per iteration, as the reference's loop prints them (ssl.py:672-677), and the
    same scores as the silent fit."""
    ti = moons["train_ind"]; tl = moons["labels"][ti]
    m = gl.ssl.poisson(moons.csr("W"), solver="gradient_descent", min_iter=7, max_iter=7)
    u_quiet = m.fit(ti, tl)
    capsys.readouterr()
    u = gl.ssl.poisson(moons.csr("W"), solver="gradient_descent", min_iter=7, max_iter=7).fit(ti, tl, all_labels=moons["labels"])
    lines = [l for l in capsys.readouterr().out.splitlines() if "Accuracy" in l]
    assert len(lines) == 7 and lines[0].startswith("1,Accuracy = ") and lines[-1].startswith("7,Accuracy = ")
    assert rel_err(u, u_quiet) <= 1e-6


def test_two_moons_fixed_T_and_directed(gl, moons):
    ti = moons["train_ind"]; tl = moons["labels"][ti]
    u = gl.ssl.poisson(moons.csr("W"), solver="gradient_descent", min_iter=80, max_iter=80).fit(ti, tl)
    assert rel_err(u, moons["u_gd_T80"]) <= TOL
    md = gl.ssl.poisson(moons.csr("Wd"), solver="gradient_descent")      # examples/poisson_directed.py
    ud = md.fit(ti, tl)
    _, T_ref = orc.poisson_gd(moons.csr("Wd"), ti, tl, return_iters=True)
    assert md.iterations == T_ref
    assert rel_err(ud, moons["u_gd_directed"]) <= TOL
    assert np.array_equal(md.predict(), moons["p_gd_directed"])


@pytest.mark.parametrize("T", [50, 200])
def test_blobs_ten_classes(gl, blobs, T):
    tb = blobs["train_ind"]
    m = gl.ssl.poisson(blobs.csr("W"), solver="gradient_descent", min_iter=T, max_iter=T)
    u = m.fit(tb, blobs["labels"][tb])
    assert m.iterations == T
    assert rel_err(u, blobs["u_gd_T%d" % T]) <= TOL
    assert np.array_equal(m.predict(), blobs["p_gd_T%d" % T])


def test_fit_predict_accuracy(gl, moons):
    ti = moons["train_ind"]
    pred = gl.ssl.poisson(moons.csr("W"), solver="gradient_descent").fit_predict(ti, moons["labels"][ti])
    assert gl.ssl.ssl_accuracy(pred, moons["labels"], ti) > 95


def test_unsorted_and_diagonal_input(gl, moons):
    """W with explicit diagonal entries and unsorted column indices gives the same answer (ssl.py:615-616)."""
    W = moons.csr("W")
    ti = moons["train_ind"]; tl = moons["labels"][ti]
    Wd = sparse.csr_matrix(W + sparse.identity(500) * 3.0)
    perm = np.random.default_rng(0)
    for i in range(500):                                   # shuffle the stored order inside each row
        s, e = Wd.indptr[i], Wd.indptr[i + 1]
        p = perm.permutation(e - s)
        Wd.indices[s:e] = Wd.indices[s:e][p]; Wd.data[s:e] = Wd.data[s:e][p]
    Wd.has_sorted_indices = False
    g = gl.graph(W); g.weight_matrix = Wd
    u = gl.ssl.poisson(g, solver="gradient_descent", min_iter=80, max_iter=80).fit(ti, tl)
    assert rel_err(u, moons["u_gd_T80"]) <= TOL


def test_isolated_node_propagates_nan_like_reference(gl):
    W = random_knn_graph(200, 5, seed=1).tolil()
    W[7, :] = 0; W[:, 7] = 0
    W = sparse.csr_matrix(W); W.eliminate_zeros()
    ti = np.array([0, 1, 2, 3]); tl = np.array([0, 1, 0, 1])
    with np.errstate(all="ignore"):
        u_ref = orc.poisson_gd(W, ti, tl, min_iter=20, max_iter=20)
    u = gl.ssl.poisson(W, solver="gradient_descent", min_iter=20, max_iter=20).fit(ti, tl)
    assert np.array_equal(np.isnan(u), np.isnan(u_ref))
    ok = ~np.isnan(u_ref)
    assert np.max(np.abs(u[ok] - u_ref[ok])) <= TOL * np.max(np.abs(u_ref[ok]))


def test_bad_sizes_raise(gl):
    from graphlearning_b200 import _lib
    lib = _lib.load()
    assert lib.glb_poisson_gd_host(None, None, None, 0, 0, None, 2, None, 0, 1, 1, None, None, None) == -1


# ---- device-resident building blocks ------------------------------------------------------------------
def test_degree_transpose_scale_bit_exact(dev, blobs):
    W = random_knn_graph(3000, 7, seed=3, symmetric=False)
    W = sparse.csr_matrix(W + sparse.identity(3000) * 0.5)        # with a diagonal to be skipped
    d = dev.DeviceCSR.from_scipy(W)
    W0 = sparse.csr_matrix(W - sparse.spdiags(W.diagonal(), 0, 3000, 3000)); W0.eliminate_zeros()
    assert np.array_equal(d.degree(skip_diagonal=False).cpu().numpy(), W * np.ones(3000))
    assert np.allclose(d.degree(skip_diagonal=True).cpu().numpy(), W0 * np.ones(3000), rtol=1e-15, atol=0)
    Wt = d.transpose().to_scipy()
    ref = sparse.csr_matrix(W.T); ref.sort_indices()
    assert np.array_equal(Wt.indptr, ref.indptr) and np.array_equal(Wt.indices, ref.indices)
    assert np.array_equal(Wt.data, ref.data)
    op = dev.PoissonOperator(W)
    s = orc.poisson_gd_setup(W, np.array([0, 1]), np.array([0, 1]))
    P = sparse.csr_matrix((op.P_val.cpu().numpy()[:op.nnz], op.col.cpu().numpy(), op.rowptr.cpu().numpy()),
                          shape=(3000, 3000))
    P.eliminate_zeros()
    Pref = s["P"].copy(); Pref.sort_indices(); Pref.eliminate_zeros()
    assert np.array_equal(P.indices, Pref.indices)
    assert np.allclose(P.data, Pref.data, rtol=1e-7, atol=0)          # fp32 rounding of an fp64 product
    RW = sparse.csr_matrix((op.RW_val.cpu().numpy()[:op.nnz], op.col.cpu().numpy(), op.rowptr.cpu().numpy()),
                           shape=(3000, 3000))
    RW.eliminate_zeros()
    RWref = s["RW"].copy(); RWref.sort_indices(); RWref.eliminate_zeros()
    assert np.allclose(RW.data, RWref.data, rtol=1e-15, atol=0)


@pytest.mark.parametrize("kind", ["dataflow", "barrier", "step"])
@pytest.mark.parametrize("c", [1, 2, 3, 4, 7, 10, 16, 17, 40, 96, 100, 130])
def test_every_kernel_every_width(dev, c, kind):
    """All three iterate kernels against the plain-C fp64 oracle for every lane mapping."""
    import torch
    W = random_knn_graph(2500, 9, seed=c)
    if kind == "dataflow" and c > 96:
        with pytest.raises(dev._lib.GlbError, match="not applicable"):
            dev.PoissonOperator(W, kind=kind).plan(c)
        return
    op = dev.PoissonOperator(W, kind=kind)
    rng = np.random.default_rng(c)
    Db64 = rng.normal(size=(2500, c)) * (rng.random((2500, 1)) < 0.05)     # sparse-ish dense source
    s = orc.poisson_gd_setup(W, np.array([0]), np.array([0]))
    Db = op.pack(Db64)
    assert op.kind(c) == kind
    assert Db.shape[1] == op.ld(c) >= c
    if kind != "dataflow":
        assert op.ld(c) == dev._lib.padded_ld(c)
    assert np.array_equal(op.unpack(Db, c).cpu().numpy(), Db64.astype(np.float32).astype(np.float64))   # pack/unpack round trip
    T = 25
    ref = c_oracle.poisson_iterate(s["P"], Db64, T)
    # one launch per iteration (the step kernel reads either layout)
    u_in = torch.zeros_like(Db); u_out = torch.zeros_like(Db)
    for _ in range(T):
        op.step(Db, u_in, u_out)
        u_in, u_out = u_out, u_in
    got_step = op.unpack(u_in, c).cpu().numpy()
    assert rel_err(got_step, ref) <= TOL
    # planned iterate
    u, launches = op.iterate(Db, T)
    got = op.unpack(u, c).cpu().numpy()
    assert rel_err(got, ref) <= TOL
    assert launches == {"dataflow": 2, "barrier": 1, "step": T}[kind]
    if kind == "dataflow":
        assert np.array_equal(got, got_step)             # same fp32 operations in the same order: bit-identical
    else:
        assert rel_err(got, got_step) <= 2e-6            # barrier kernel sums long rows in a different (fixed) order
    u_again, _ = op.iterate(Db, T)
    assert np.array_equal(op.unpack(u_again, c).cpu().numpy(), got)      # run-to-run deterministic


@pytest.mark.parametrize("kind", ["dataflow", "barrier", "step"])
def test_iterate_T_zero_and_odd_even(dev, kind):
    import torch
    W = random_knn_graph(1000, 6, seed=5)
    op = dev.PoissonOperator(W, kind=kind)
    Db = op.pack(np.random.default_rng(0).normal(size=(1000, 10)))
    u0, _ = op.iterate(Db, 0)
    assert float(op.unpack(u0, 10).abs().max()) == 0.0
    u3, _ = op.iterate(Db, 3)
    u4, _ = op.iterate(Db, 4)
    a = torch.zeros_like(Db)
    op.step(Db, u3, a)
    assert torch.equal(op.unpack(a, 10), op.unpack(u4, 10))
    # a warm start: T1 + T2 iterations in two calls == T1 + T2 in one
    u7, _ = op.iterate(Db, 7)
    u34, _ = op.iterate(Db, 3, u0=u4.clone())
    assert torch.equal(op.unpack(u34, 10), op.unpack(u7, 10))


def test_kernel_choice(dev):
    """auto: symmetric pattern -> dataflow; directed -> barrier; explicit request that does not apply -> error."""
    Ws = random_knn_graph(3000, 8, seed=2)
    Wd = random_knn_graph(3000, 8, seed=2, symmetric=False)
    assert dev.PoissonOperator(Ws).kind(10) in ("dataflow", "barrier")   # auto times both on a trial run
    assert dev.PoissonOperator(Ws, kind="dataflow").kind(10) == "dataflow"
    assert dev.PoissonOperator(Wd).kind(10) == "barrier"
    assert dev.PoissonOperator(Ws).kind(200) == "barrier"         # too wide for the flagged layout
    with pytest.raises(dev._lib.GlbError, match="not applicable"):
        dev.PoissonOperator(Wd, kind="dataflow").plan(10)
    op = dev.PoissonOperator(Ws, kind="dataflow")
    assert 0.5 < op.fill(10) <= 1.0                               # slice widths are padded to the 8-entry octets of the pipelined stream


@pytest.mark.parametrize("c", [5, 10, 40])
def test_dataflow_ragged_rows(dev, c):
    """Empty rows, hub rows split over several warps (400 and 1700 nonzeros), rows shorter than the unroll: the
    sliced-ELL slabs and the shared-memory partial sums stay exact for every lane mapping."""
    rng = np.random.default_rng(7)
    n = 2500
    W = random_knn_graph(n, 3, seed=9).tolil()
    W[5, :] = 0; W[:, 5] = 0                                      # isolated node
    for hub_node, deg in ((11, 400), (12, 1700), (2499, 40)):
        hub = rng.choice(n, deg, replace=False)
        hub = hub[(hub != hub_node) & (hub != 5)]
        W[hub_node, hub] = 0.5; W[hub, hub_node] = 0.5
    W = sparse.csr_matrix(W); W.eliminate_zeros()
    op = dev.PoissonOperator(W, kind="dataflow")
    # the operator's own fp32 P (row 5 and column 5 are empty, so the zero degree injects nothing)
    P = sparse.csr_matrix((op.P_val.cpu().numpy()[:op.nnz].astype(np.float64), op.col.cpu().numpy(),
                           op.rowptr.cpu().numpy()), shape=(n, n))
    assert np.isfinite(P.data).all() and np.diff(P.indptr).max() > 1500 and np.diff(P.indptr).min() == 0
    Db64 = rng.normal(size=(n, c))
    ref = c_oracle.poisson_iterate(P, Db64, 12)
    Db = op.pack(Db64)
    got = op.unpack(op.iterate(Db, 12)[0], c).cpu().numpy()
    assert rel_err(got, ref) <= TOL
    again = op.unpack(op.iterate(Db, 12)[0], c).cpu().numpy()
    assert np.array_equal(got, again)                             # fixed summation order: run-to-run identical


def test_mixing_T_matches_reference_rule(dev, blobs, moons):
    for g, key in ((moons, "W"), (moons, "Wd"), (blobs, "W")):
        W = g.csr(key); ti = g["train_ind"]
        _, T_ref = orc.poisson_gd(W, ti, g["labels"][ti], return_iters=True)
        op = dev.PoissonOperator(W)
        assert op.mixing_T(ti, 50, 1000) == T_ref
        assert op.mixing_T(ti, 0, 7) == 7
        assert op.mixing_T(ti, 5, 5) == 5


def test_mixing_T_beyond_one_grid_pass(dev):
    """More rows than one pass of the persistent mixing kernel holds in registers (75 776 on 148 SMs), rule deciding at a T
    that is neither min_iter nor max_iter, min_iter = 0 (err_0 is read), odd batch boundaries."""
    n = 90000
    W = random_knn_graph(n, 6, seed=5)
    ti = np.arange(0, n, 9000)
    s = orc.poisson_gd_setup(W, ti, np.arange(len(ti)) % 2)
    v, errs = s["v"], []
    for _ in range(80):
        errs.append(np.max(np.absolute(v - s["vinf"])))
        v = s["RW"] * v
    def rule(lo, hi):
        T = 0
        while (T < lo or errs[T] > 1 / n) and T < hi:
            T += 1
        return T
    assert 2 < rule(0, 80) < 80                       # the rule itself decides on this graph
    op = dev.PoissonOperator(W)
    for lo, hi in ((0, 80), (3, 80), (0, 5), (70, 80)):
        assert op.mixing_T(ti, lo, hi) == rule(lo, hi)


# ---- north-star size (70k nodes, k=10, 10 classes) ----------------------------------------------------
@pytest.fixture(scope="module")
def big():
    W = random_knn_graph(70000, 10, seed=0)
    labels = np.random.default_rng(1).integers(0, 10, 70000)
    ti = orc.one_per_class(labels, rate=1, seed=0)
    return W, labels, ti


def test_full_size_against_c_oracle(dev, big):
    W, labels, ti = big
    s = orc.poisson_gd_setup(W, ti, labels[ti])
    ref = c_oracle.poisson_iterate(s["P"], s["Db"], 60)
    for kind, nl in (("dataflow", 2), ("barrier", 1)):
        op = dev.PoissonOperator(W, kind=kind)
        Db = op.source_to_Db(orc.poisson_source(70000, ti, labels[ti])[0])
        assert op.kind(10) == kind
        u, launches = op.iterate(Db, 60)
        assert launches == nl
        assert rel_err(op.unpack(u, 10).cpu().numpy(), ref) <= TOL


def test_full_size_properties(dev, big):
    """Size-independent properties: linearity in the source, zero class-sum invariant, step/persistent equality."""
    import torch
    W, labels, ti = big
    op = dev.PoissonOperator(W, kind="dataflow")
    rng = np.random.default_rng(2)
    A = op.pack(rng.normal(size=(70000, 10))); B = op.pack(rng.normal(size=(70000, 10)))
    T = 30
    assert op.kind(10) == "dataflow"
    uA = op.unpack(op.iterate(A, T)[0], 10); uB = op.unpack(op.iterate(B, T)[0], 10)
    uAB = op.unpack(op.iterate(A * 2 - B, T)[0], 10)
    lin = 2 * uA - uB
    assert float((uAB - lin).abs().max() / lin.abs().max()) <= 2e-5
    src = orc.poisson_source(70000, ti, labels[ti])[0]
    u = op.unpack(op.iterate(op.source_to_Db(src), 1000)[0], 10)
    rowsum = u.sum(1).abs().max()
    assert float(rowsum) <= 1e-5 * float(u.abs().max())          # columns of the source sum to zero per row
    a = torch.zeros_like(A); b = torch.zeros_like(A)
    for _ in range(T):
        op.step(A, a, b); a, b = b, a
    # same arithmetic as T single steps except for rows longer than 32 (summed by a whole warp in a fixed order)
    assert float((op.unpack(a, 10) - uA).abs().max() / uA.abs().max()) <= 2e-6


def test_full_size_through_host_api(gl, big):
    W, labels, ti = big
    m = gl.ssl.poisson(W, solver="gradient_descent", min_iter=40, max_iter=40)
    u = m.fit(ti, labels[ti])
    s = orc.poisson_gd_setup(W, ti, labels[ti])
    ref = c_oracle.poisson_iterate(s["P"], s["Db"], 40)
    assert rel_err(u, ref) <= TOL
    # labels identical except where the reference's own top-2 margin is below the score tolerance
    pred, pref = m.predict(), orc.predict(ref)
    srt = np.sort(ref, axis=1)
    margin = (srt[:, -1] - srt[:, -2]) / np.max(np.abs(ref))
    assert np.all((pred == pref) | (margin < 2 * TOL))
    assert np.mean(pred == pref) > 0.999


def test_sparse_source_fit_is_the_dense_fit(gl, blobs):
    """glb_poisson_graph_fit_rows (labelled rows only over PCIe, pinned result) = glb_poisson_graph_fit on the dense source,
    bit for bit, with numpy's assignment semantics: negative indices, of a repeated index the last row stays."""
    import gc
    W = blobs.csr("W")
    n = W.shape[0]
    h = gl.graph(W).poisson_handle()
    rng = np.random.default_rng(3)
    ti = rng.choice(n, 40, replace=False)
    ti[5] = ti[30]                                  # repeated node: the row of position 30 must win
    ti[7] -= n                                      # negative index
    rows = rng.standard_normal((40, 6))
    src = np.zeros((n, 6))
    src[ti] = rows
    for lo, hi in ((30, 30), (20, 200)):
        u_dense, T_dense, _ = h.fit(src, ti, lo, hi)
        u_rows, T_rows, nl = h.fit_rows(ti, rows, ti, lo, hi)
        assert T_rows == T_dense and nl > 0
        assert np.array_equal(u_rows, u_dense)
    # results move to page-locked memory once the pool's background thread has pinned buffers of this size; a buffer
    # returns to the pool when the array and its views are gone
    from graphlearning_b200 import device
    count = lambda: sum(len(v) for v in device.pinned.idle.values())
    device.pinned.wait()
    idle = count()
    assert idle >= 1
    u_pin, _, _ = h.fit_rows(ti, rows, ti, lo, hi)
    assert count() == idle - 1 and np.array_equal(u_pin, u_dense)
    view = u_pin[:, 1]
    del u_pin
    gc.collect()
    assert count() == idle - 1 and np.array_equal(view, u_dense[:, 1])
    del view
    gc.collect()
    assert count() == idle
    with pytest.raises(Exception):
        h.fit_rows(np.array([n]), rows[:1], np.array([0]), 5, 5)      # out of range like numpy's IndexError
    u0, _, _ = h.fit_rows(np.zeros(0, dtype=np.int64), np.zeros((0, 6)), np.zeros(0, dtype=np.int64), 5, 5)
    assert np.array_equal(u0, np.zeros((n, 6)))
