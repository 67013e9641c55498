"""Test double of the device block kernels (spectral.cu) with the same call contract as
graphlearning_b200.spectral.BlockOps, in numpy.  Lets the CPU suite exercise the HOST logic of the spectral solver
(filter recurrence, degree selection, orthonormalisation, Rayleigh-Ritz); the -m gpu tests run the real kernels."""
import numpy as np
from scipy import sparse

from graphlearning_b200 import spectral


class _Arr(np.ndarray):
    def cpu(self):
        return self

    def numpy(self):
        return np.asarray(self)


class NumpyOps(spectral.BlockOps):
    def __init__(self, A):
        A = sparse.csr_matrix(A)
        self.n = A.shape[0]
        self.Ah, self.Ath = A, A.T.tocsr()
        self.launches = 0

    def new(self, c):
        return np.zeros((self.n, spectral._even(c))).view(_Arr)

    def upload(self, X):
        t = self.new(X.shape[1])
        t[:, : X.shape[1]] = X
        return t

    def spmm(self, X, c, out=None, transpose=False, alpha=1.0, Y1=None, beta=0.0, bcol=None, Y2=None, gamma=0.0):
        M = self.Ath if transpose else self.Ah
        r = alpha * (M @ np.asarray(X[:, :c]))
        if Y1 is not None:
            r = r + beta * np.asarray(Y1[:, :c]) * (1.0 if bcol is None else np.asarray(bcol)[None, :])
        if Y2 is not None:
            r = r + gamma * np.asarray(Y2[:, :c])
        if out is None:
            out = self.new(c)
        out[:, :c] = r
        out[:, c:] = 0
        self.launches += 1
        return out

    def set_columns(self, X, idx, R):
        X[:, idx] = R

    def gram(self, X, c1, Y, c2):
        self.launches += 2
        return np.asarray(X[:, :c1]).T @ np.asarray(Y[:, :c2])

    def right_mul(self, X, c1, S, out=None):
        S = np.asarray(S)
        if out is None:
            out = self.new(S.shape[1])
        out[:, : S.shape[1]] = np.asarray(X[:, :c1]) @ S
        out[:, S.shape[1]:] = 0
        self.launches += 1
        return out
