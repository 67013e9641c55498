"""Leading singular pairs of a sparse matrix on the GPU: the solver behind graph.eigen_decomp and
utils.randomized_svd (reference graphlearning/graph.py:623-806, graphlearning/utils.py:576-642).

The reference hands A = D^-1/2 W D^-1/2 (or 2 max(deg) I - L) to ARPACK `svds` (method='exact', graph.py:734,756) or
to a randomized SVD (method='lowrank', utils.py:611-639).  Here both run on block operations of libglb200.so
(spectral.cu: fp64 CSR SpMM with a fused three-term recurrence, Gram matrix, right-multiplication by a small matrix):

  svd_topk        block Chebyshev-filtered subspace iteration on M = A^T A with Rayleigh-Ritz, to ARPACK accuracy;
                  the c x c dense eigenproblems (c = k + guard vectors <= 256) are solved on the host with LAPACK,
                  exactly the kind of small dense step the reference does in numpy.
  randomized_svd  the reference's algorithm step by step (same Gaussian test matrix from numpy's global RNG, same
                  power iteration Y <- A (A^T Y)), with the QR / SVD of the tall-skinny factors done through Gram
                  matrices on the device.

torch only owns the device memory.  There is no CPU fallback: every n-sized product runs through the C-ABI.
"""
from __future__ import annotations

import ctypes

import numpy as np
from scipy import sparse

from . import _lib
from .device import DeviceCSR, _torch, cur_stream, ptr


def _even(c):
    return (int(c) + 1) & ~1


class BlockOps:
    """A (CSR fp64) and A^T resident in HBM plus the three block kernels on n x ld fp64 row-major tensors."""

    def __init__(self, A):
        torch = _torch()
        A = sparse.csr_matrix(A)
        if A.shape[0] != A.shape[1]:
            raise ValueError("square matrices only")
        A.sort_indices()
        self.n = A.shape[0]
        self.A = DeviceCSR.from_scipy(A)
        At = A.T.tocsr()
        At.sort_indices()
        symmetric = (np.array_equal(A.indptr, At.indptr) and np.array_equal(A.indices, At.indices)
                     and np.array_equal(A.data, At.data))
        self.At = self.A if symmetric else self.A.transpose()
        self.symmetric = symmetric
        self.launches = 0
        self._work = None
        self.torch = torch

    def new(self, c):
        return self.torch.zeros((self.n, _even(c)), dtype=self.torch.float64, device="cuda")

    def upload(self, X):
        X = np.asarray(X, dtype=np.float64)
        t = self.new(X.shape[1])
        t[:, : X.shape[1]] = self.torch.from_numpy(np.ascontiguousarray(X)).to("cuda")
        return t

    def spmm(self, X, c, out=None, transpose=False, alpha=1.0, Y1=None, beta=0.0, bcol=None, Y2=None, gamma=0.0):
        """out = alpha * op(A) X + beta * Y1 diag(bcol) + gamma * Y2   (op = transpose or identity)."""
        M = self.At if transpose else self.A
        if out is None:
            out = self.new(c)
        bc = None
        if bcol is not None:
            bc = self.torch.from_numpy(np.ascontiguousarray(bcol, dtype=np.float64)).to("cuda")
        _lib.call("glb_spmm_f64", ptr(M.rowptr), ptr(M.col), ptr(M.val), self.n, ptr(X), X.shape[1], ptr(out), out.shape[1],
                  int(c), float(alpha), ptr(Y1), Y1.shape[1] if Y1 is not None else 0, float(beta), ptr(bc), ptr(Y2),
                  Y2.shape[1] if Y2 is not None else 0, float(gamma), cur_stream())
        self.launches += 1
        return out

    def gram(self, X, c1, Y, c2):
        """X^T Y as a (c1, c2) numpy array."""
        torch = self.torch
        wb = int(_lib.load().glb_gram_work_bytes(int(c1), int(c2)))
        if self._work is None or self._work.numel() < wb:
            self._work = torch.empty(wb, dtype=torch.uint8, device="cuda")
        G = torch.empty((c1, c2), dtype=torch.float64, device="cuda")
        _lib.call("glb_gram_f64", ptr(X), X.shape[1], int(c1), ptr(Y), Y.shape[1], int(c2), self.n, ptr(G), ptr(self._work),
                  wb, cur_stream())
        self.launches += 2
        return G.cpu().numpy()

    def right_mul(self, X, c1, S, out=None):
        """X[:, :c1] @ S for a small host matrix S (c1, c2)."""
        S = np.ascontiguousarray(S, dtype=np.float64)
        c2 = S.shape[1]
        if out is None:
            out = self.new(c2)
        Sd = self.torch.from_numpy(S).to("cuda")
        _lib.call("glb_right_mul_f64", ptr(X), X.shape[1], self.n, int(c1), ptr(Sd), int(c2), ptr(out), out.shape[1],
                  cur_stream())
        self.launches += 1
        return out

    def set_columns(self, X, idx, R):
        X[:, idx] = self.torch.from_numpy(np.ascontiguousarray(R, dtype=np.float64)).to("cuda")

    def orthonormalize(self, X, c, rng=None):
        """Columns of X[:, :c] -> an orthonormal basis (column-scaled Cholesky QR, repeated until the Gram matrix is the
        identity to 1e-13).  When the Gram matrix is numerically singular a rank-revealing eigen-decomposition keeps the
        significant directions; the lost ones are replaced by random vectors when `rng` is given (subspace iteration)
        or left as zero columns (randomized SVD, where they carry zero singular values)."""
        for it in range(8):
            G = self.gram(X, c, X, c)
            d = np.sqrt(np.maximum(np.diag(G), 0.0))
            live = d > 0
            d[~live] = 1.0
            Gs = G / np.outer(d, d)
            off = np.max(np.abs(Gs[np.ix_(live, live)] - np.eye(int(live.sum())))) if live.any() else 0.0
            if it > 0 and off < 1e-13 and (live.all() or rng is None):
                break
            lost = np.zeros(0, dtype=np.int64)
            try:
                if not live.all():
                    raise np.linalg.LinAlgError
                L = np.linalg.cholesky(Gs)
                S = np.linalg.solve(L, np.diag(1.0 / d)).T            # diag(1/d) L^-T
            except np.linalg.LinAlgError:
                Gs[~live, :] = 0.0
                Gs[:, ~live] = 0.0
                w, V = np.linalg.eigh(Gs)
                keep = w > 1e-13 * max(w.max(), 1e-300)
                nk = int(keep.sum())
                S = np.zeros((c, c))
                S[:, :nk] = (V[:, keep] / np.sqrt(w[keep])) / d[:, None]
                lost = np.arange(nk, c)
            X = self.right_mul(X, c, S)
            if len(lost) and rng is not None:
                self.set_columns(X, lost, rng.standard_normal((self.n, len(lost))))
        return X


def svd_topk(A, k, tol=0.0, guard=None, degree=None, max_outer=200, seed=0, return_info=False):
    """k largest singular values of the square sparse matrix A with left singular vectors: (u (n,k), s (k,) descending).

    Block Chebyshev-filtered subspace iteration on M = A^T A: the filter damps [0, smallest Ritz value of the
    block] and is applied through the fused SpMM recurrence; every outer step ends with a Rayleigh-Ritz projection and
    a residual check ||M v - s^2 v|| <= rtol * s_max^2 on the k wanted pairs (rtol = max(tol, 2e-12), i.e. ARPACK's
    machine-precision default tol=0 of the reference call)."""
    ops = A if isinstance(A, BlockOps) else BlockOps(A)
    n = ops.n
    k = int(k)
    if k < 1 or k >= n:
        raise ValueError("k must satisfy 1 <= k < n")
    nb = min(n, k + (max(8, k // 4) if guard is None else int(guard)))
    nb = min(nb, 256)
    if k > nb:
        raise ValueError("k too large for the block solver (<= 248)")
    rtol = max(float(tol), 2e-12)
    rng = np.random.default_rng(seed)
    X = ops.orthonormalize(ops.upload(rng.standard_normal((n, nb))), nb, rng)
    T = ops.new(nb)
    m_max = 40 if degree is None else int(degree)
    info = {"outer": 0, "spmm": 0, "residual": None}
    w = None
    best = np.inf
    for outer in range(max_outer):
        ops.spmm(X, nb, out=T)                                             # T = A X
        H = ops.gram(T, nb, T, nb)                                         # X^T M X
        H = (H + H.T) / 2
        w, Q = np.linalg.eigh(H)
        w, Q = w[::-1].copy(), Q[:, ::-1].copy()
        X = ops.right_mul(X, nb, Q)
        T = ops.right_mul(T, nb, Q)
        R = ops.spmm(T, nb, transpose=True, Y1=X, beta=-1.0, bcol=w)      # A^T (A X) - X diag(w)
        rn = np.sqrt(np.maximum(np.diag(ops.gram(R, nb, R, nb)), 0.0))
        info["spmm"] += 2
        res = float(np.max(rn[:k]) / max(w[0], 1e-300))
        info.update(outer=outer, residual=res, converged=bool(res <= rtol), rtol=rtol)
        if res <= rtol:
            break
        if outer > 12 and res > 0.5 * best and res < 1e-9:                 # stagnation at the rounding floor
            info["stagnated"] = True
            break
        best = min(best, res)
        # Chebyshev filter: damp [0, b], amplify above; scaled three-term recurrence (Zhou & Saad)
        a, b, a0 = 0.0, float(max(w[-1], 1e-300)), float(w[0])
        if not a0 > b * (1 + 1e-12):
            a0 = b * 1.01 + 1e-300
        e, cen = (b - a) / 2, (b + a) / 2
        # degree: as high as the dynamic range allows.  The filter multiplies the top pair by T_m(x_top) and the k-th
        # wanted pair by T_m(x_k); beyond a ratio of ~1e7 the block collapses onto the dominant directions faster than
        # fp64 can orthogonalise it again.
        gap = np.arccosh((a0 - cen) / e) - np.arccosh(max((float(w[k - 1]) - cen) / e, 1.0))
        m = m_max if gap <= 0 else int(min(m_max, max(4, np.floor(np.log(1e7) / gap))))
        if outer < 2:
            m = min(m, 4 << outer)          # the first Ritz values badly underestimate the top of the spectrum
        sigma = e / (a0 - cen)
        tau = 2.0 / sigma
        ops.spmm(X, nb, out=T)
        Y = ops.spmm(T, nb, transpose=True, alpha=sigma / e, Y1=X, beta=-cen * sigma / e)
        for _ in range(2, m + 1):
            sn = 1.0 / (tau - sigma)
            ops.spmm(Y, nb, out=T)
            # X <- (2 sn / e) (M Y - cen Y) - sigma sn X   (written over the old X, which is only read row-locally)
            ops.spmm(T, nb, out=X, transpose=True, alpha=2 * sn / e, Y1=Y, beta=-cen * 2 * sn / e, Y2=X, gamma=-sigma * sn)
            X, Y = Y, X
            sigma = sn
        info["spmm"] += 2 * m
        info.setdefault("degrees", []).append(m)
        X = ops.orthonormalize(Y, nb, rng)
    s = np.sqrt(np.maximum(w[:k], 0.0))
    U = ops.spmm(X, nb)                                                    # left singular vectors u = A v / s
    u = U[:, :k].cpu().numpy() / np.where(s > 0, s, 1.0)
    info["launches"] = ops.launches
    if return_info:
        return u, s, info
    return u, s


def randomized_svd(A, k=10, c=None, q=1, return_info=False):
    """Randomized SVD, mirror of reference graphlearning/utils.py:576-642 (same steps, same use of numpy's global
    RNG for the test matrix); returns (u (n,k), s (k,), vt (k,n))."""
    if c is None:
        c = 2 * k
    ops = A if isinstance(A, BlockOps) else BlockOps(A)
    n = ops.n
    c = int(c)
    if c > 256:
        raise ValueError("c <= 256 on the B200 backend")
    Omega = np.random.randn(n, c)                                          # utils.py:614
    Y = ops.spmm(ops.upload(Omega), c)                                     # Y = A Omega            :615
    T = ops.new(c)
    for _ in range(int(q)):                                                # Y = A (A^T Y)          :616-617
        ops.spmm(Y, c, out=T, transpose=True)
        ops.spmm(T, c, out=Y)
    Q = ops.orthonormalize(Y, c)                                           # Q, R = qr(Y)           :620
    Z = ops.spmm(Q, c, transpose=True)                                     # B = Q^T A  <=>  Z = A^T Q = B^T   :623
    G = ops.gram(Z, c, Z, c)                                               # B B^T
    w, Ub = np.linalg.eigh((G + G.T) / 2)
    w, Ub = w[::-1], Ub[:, ::-1]                                           # sorted from largest to smallest  :627-630
    s = np.sqrt(np.maximum(w, 0.0))
    kk = min(int(k), c)
    u = ops.right_mul(Q, c, Ub[:, :kk])[:, :kk].cpu().numpy()              # u = Q u_B              :625
    sk = s[:kk]
    v = ops.right_mul(Z, c, Ub[:, :kk] / np.where(sk > 0, sk, 1.0))[:, :kk].cpu().numpy()
    if return_info:
        return u, sk.copy(), v.T.copy(), {"launches": ops.launches}
    return u, sk.copy(), v.T.copy()
