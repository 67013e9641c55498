"""ctypes doorway onto oracle/_build/liboracle.so (the plain-C restatement, oracle_kernels.c)
and oracle/_ref/liblp_ref.so (the reference's own c_code/lp_iterate.cpp compiled in place).

TEST INFRASTRUCTURE ONLY - see gl_oracle.py header for who may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "liblp_ref.so")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile the restatement (and, when /root/reference is present, oracle/_ref)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(
            os.path.join(_HERE, "oracle_kernels.c")):
        subprocess.run(["make", "-C", _HERE, "_build/liboracle.so"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/c_code") and (force or not os.path.exists(_REF)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB)
        _lib.orc_mixing_iterate.restype = ctypes.c_double
    return _lib


def ref_available():
    return os.path.exists(_REF)


def ref():
    global _ref
    if _ref is None:
        if not ref_available():
            build()
        _ref = ctypes.CDLL(_REF)
    return _ref


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _csr(A):
    return (np.ascontiguousarray(A.indptr, dtype=np.int32), np.ascontiguousarray(A.indices, dtype=np.int32),
            np.ascontiguousarray(A.data, dtype=np.float64))


def csr_matvecs(A, X):
    """Y = A @ X restated row by row (scipy _sparsetools csr_matvecs)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    p, j, x = _csr(A)
    Y = np.zeros((A.shape[0], X.shape[1]))
    lib().orc_csr_matvecs(ctypes.c_int(A.shape[0]), ctypes.c_int(X.shape[1]), _i(p), _i(j), _d(x), _d(X), _d(Y))
    return Y


def poisson_iterate(P, Db, T, u0=None):
    """T steps of u <- Db + P*u (graphlearning/ssl.py:667-668), fp64, one thread."""
    Db = np.ascontiguousarray(Db, dtype=np.float64)
    u = np.zeros_like(Db) if u0 is None else np.ascontiguousarray(u0, dtype=np.float64).copy()
    p, j, x = _csr(P)
    rc = lib().orc_poisson_iterate(ctypes.c_int(P.shape[0]), ctypes.c_int(Db.shape[1]), _i(p), _i(j), _d(x),
                                   _d(Db), _d(u), ctypes.c_int(int(T)))
    if rc != 0:
        raise MemoryError
    return u


def mixing_iterate(RW, v, vinf, T):
    v = np.ascontiguousarray(v, dtype=np.float64).copy()
    vinf = np.ascontiguousarray(vinf, dtype=np.float64)
    p, j, x = _csr(RW)
    err = lib().orc_mixing_iterate(ctypes.c_int(RW.shape[0]), _i(p), _i(j), _d(x), _d(vinf), _d(v),
                                   ctypes.c_int(int(T)))
    return v, err


def _pad1(a, dtype):
    """The reference scans `J[j]==i & j<M` reading J[M] before the bound check
    (c_code/lp_iterate.cpp:51,141): hand it arrays with one spare element."""
    out = np.empty(len(a) + 1, dtype=dtype)
    out[:-1] = a
    out[-1] = -1
    return out


def lp_iterate(uu, ul, I, J, W, ind, val, p, T, tol, use_ref=False):
    """Returns (uu, ul[, sweeps]) as the caller's buffers hold them on return."""
    uu = np.ascontiguousarray(uu, dtype=np.float64).copy()
    ul = np.ascontiguousarray(ul, dtype=np.float64).copy()
    I_ = _pad1(I, np.int32); J_ = _pad1(J, np.int32); W_ = _pad1(W, np.float64)
    ind = np.ascontiguousarray(ind, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
    n, M, m = len(uu), len(I), len(ind)
    if use_ref:
        ref().ref_lp_iterate(_d(uu), _d(ul), _i(I_), _i(J_), _d(W_), _i(ind), _d(val), ctypes.c_double(p),
                             ctypes.c_int(int(T)), ctypes.c_double(tol), ctypes.c_int(n), ctypes.c_int(M),
                             ctypes.c_int(m))
        return uu, ul
    sweeps = lib().orc_lp_iterate(_d(uu), _d(ul), _i(I_), _i(J_), _d(W_), _i(ind), _d(val), ctypes.c_double(p),
                                  ctypes.c_int(int(T)), ctypes.c_double(tol), ctypes.c_int(n), ctypes.c_int(M),
                                  ctypes.c_int(m))
    return uu, ul, sweeps


def lip_iterate(u, I, J, W, ind, val, T, tol, alpha, beta, use_ref=False):
    u = np.ascontiguousarray(u, dtype=np.float64).copy()
    I_ = _pad1(I, np.int32); J_ = _pad1(J, np.int32); W_ = _pad1(W, np.float64)
    ind = np.ascontiguousarray(ind, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
    n, M, m = len(u), len(I), len(ind)
    if use_ref:
        ref().ref_lip_iterate(_d(u), _i(I_), _i(J_), _d(W_), _i(ind), _d(val), ctypes.c_int(int(T)),
                              ctypes.c_double(tol), ctypes.c_int(n), ctypes.c_int(M), ctypes.c_int(m),
                              ctypes.c_double(alpha), ctypes.c_double(beta))
        return u
    sweeps = lib().orc_lip_iterate(_d(u), _i(I_), _i(J_), _d(W_), _i(ind), _d(val), ctypes.c_int(int(T)),
                                   ctypes.c_double(tol), ctypes.c_int(n), ctypes.c_int(M), ctypes.c_int(m),
                                   ctypes.c_double(alpha), ctypes.c_double(beta))
    return u, sweeps


def lip_iterate_weighted(u, I, J, W, ind, val, T, tol, use_ref=False):
    """c_code/lp_iterate.cpp:190-259.  Returns u (reference build) or (u, sweeps) (restatement)."""
    u = np.ascontiguousarray(u, dtype=np.float64).copy()
    I_ = _pad1(I, np.int32); J_ = _pad1(J, np.int32); W_ = _pad1(W, np.float64)
    I_[-1] = 0                              # an empty LAST row makes the reference read u[I[M]]: keep that read in bounds
    ind = np.ascontiguousarray(ind, dtype=np.int32); val = np.ascontiguousarray(val, dtype=np.float64)
    n, M, m = len(u), len(I), len(ind)
    fn = ref().ref_lip_iterate_weighted if use_ref else lib().orc_lip_iterate_weighted
    sweeps = fn(_d(u), _i(I_), _i(J_), _d(W_), _i(ind), _d(val), ctypes.c_int(int(T)), ctypes.c_double(tol),
                ctypes.c_int(n), ctypes.c_int(M), ctypes.c_int(m))
    return u if use_ref else (u, sweeps)
