"""Host logic of the spectral solver (graphlearning_b200/spectral.py) on CPU, with the device kernels replaced by
the numpy test double tests/_numpy_blockops.py, against dense LAPACK and the reference-generated goldens."""
import numpy as np
from scipy import sparse

from _numpy_blockops import NumpyOps
from graphlearning_b200 import spectral


def normalized_adjacency(W):
    deg = np.asarray(W.sum(axis=1)).ravel()
    D = sparse.spdiags(deg ** -0.5, 0, W.shape[0], W.shape[0])
    return sparse.csr_matrix(D @ W @ D)


def test_svd_topk_repeated_and_clustered_spectrum(blobs):
    A = normalized_adjacency(blobs.csr("W"))                 # ten disconnected blobs: eigenvalue 1 ten times
    sv = np.sort(np.abs(np.linalg.eigvalsh(A.toarray())))[::-1]
    for k in (10, 50):
        u, s, info = spectral.svd_topk(NumpyOps(A), k, return_info=True)
        assert np.abs(s - sv[:k]).max() < 1e-10
        assert np.abs(u.T @ u - np.eye(k)).max() < 1e-9
        assert np.abs(A @ (A.T @ u) - u * s ** 2).max() < 1e-9
        assert info["residual"] < 1e-9 and max(info["degrees"]) <= 40


def test_svd_topk_nonsymmetric():
    rng = np.random.default_rng(0)
    A = sparse.random(300, 300, density=0.05, random_state=1, format="csr") + sparse.identity(300) * 0.1
    u, s = spectral.svd_topk(NumpyOps(A), 6)
    sd = np.linalg.svd(A.toarray(), compute_uv=False)
    assert np.abs(s - sd[:6]).max() < 1e-9


def test_randomized_svd_follows_reference_algorithm(moons):
    A = normalized_adjacency(moons.csr("W"))
    np.random.seed(1)
    u, s, vt = spectral.randomized_svd(NumpyOps(A), k=8, c=30, q=20)
    np.random.seed(1)
    Om = np.random.randn(500, 30)                            # utils.py:614-639 restated
    Y = A @ Om
    for _ in range(20):
        Y = A @ (A.T @ Y)
    Q, R = np.linalg.qr(Y)
    ub, sb, vtb = np.linalg.svd(Q.T @ A, full_matrices=False)
    ub = Q @ ub
    assert np.abs(s - sb[:8]).max() < 1e-9
    assert np.abs(u @ u.T - ub[:, :8] @ ub[:, :8].T).max() < 1e-6
    assert u.shape == (500, 8) and vt.shape == (8, 500)
