"""Exact kNN search on the GPU (knn.cu) behind weightmatrix.knnsearch."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib

last_stats = {}


def knnsearch_gpu(X, k, similarity="euclidean"):
    """(knn_ind (n,k) int64, knn_dist (n,k) float64) including self, ascending - the contract of reference
    graphlearning/weightmatrix.py:297-429.  Angular similarity = Euclidean on row-normalised data (:344-345)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim != 2:
        raise ValueError("X must be (n, d)")
    n, d = X.shape
    if similarity == "angular":
        X = X / np.linalg.norm(X, axis=1)[:, None]
    elif similarity != "euclidean":
        raise ValueError("Invalid choice of similarity " + str(similarity))
    X = np.ascontiguousarray(X)
    k = int(k)
    ind = np.empty((n, k), dtype=np.int64)
    dist = np.empty((n, k), dtype=np.float64)
    nl, nf = ctypes.c_int(0), ctypes.c_int(0)
    _lib.call("glb_knn_search_host", ctypes.c_void_p(X.ctypes.data), n, d, k, ctypes.c_void_p(ind.ctypes.data),
              ctypes.c_void_p(dist.ctypes.data), ctypes.byref(nl), ctypes.byref(nf))
    last_stats.update(launches=nl.value, fallback_rows=nf.value)
    return ind, dist
