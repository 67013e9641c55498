"""kNN weight matrices: mirror of reference graphlearning/weightmatrix.py:68-187 (knn) and :297-429 (knnsearch).

knnsearch routes method in {None (d>5), 'brute', 'annoy'} with euclidean/angular similarity to the exact
tiled brute-force search on the GPU (knn.cu) - a strict quality superset of the reference's approximate
annoy default.  method='kdtree' (the reference default for d<=5) stays scipy's cKDTree, exactly as in the
reference (weightmatrix.py:349-352).
"""
from __future__ import annotations

import os

import numpy as np
from scipy import sparse, spatial

# stored kNN searches live in ./knn_data/<dataset>_<metric>.npz with fields J (indices) and D (distances): the wire
# format between graph build and solve of the reference (weightmatrix.py:17, 416-427, 431-467)
knn_dir = os.path.abspath(os.path.join(os.getcwd(), "knn_data"))


# graphs below this size are assembled with scipy on the host (a device round trip costs more than it saves);
# tests set it to 0 to exercise the device path on the small goldens
_device_assembly_min_n = 2048


def _assemble_on_device(knn_ind, weights, n, k, symmetrize):
    """weightmatrix.py:166-186 through glb_knn_weights_csr_host (knn_graph.cu).  symmetrize: 0 none, 1 average,
    2 sparse_max, 3 the symgaussian rule."""
    import ctypes
    from . import _lib
    ind = np.ascontiguousarray(knn_ind, dtype=np.int64)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    cap = n * k * (2 if symmetrize else 1)
    rp = np.empty(n + 1, dtype=np.int32)
    col = np.empty(cap, dtype=np.int32)
    val = np.empty(cap, dtype=np.float64)
    nnz, nl = ctypes.c_int64(0), ctypes.c_int(0)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    _lib.call("glb_knn_weights_csr_host", p(ind), p(w), int(n), int(k), int(symmetrize), p(rp), p(col), p(val), int(cap),
              ctypes.byref(nnz), ctypes.byref(nl))
    m = nnz.value
    return sparse.csr_matrix((val[:m].copy(), col[:m].copy(), rp), shape=(n, n))


def sparse_max(A, B):
    """Elementwise max of two sparse matrices.  Reference graphlearning/utils.py:263-286."""
    I = (A + B) > 0
    IB = B > A
    IA = I - IB
    return A.multiply(IA) + B.multiply(IB)


def knnsearch(X, k, method=None, similarity="euclidean", dataset=None, metric="raw"):
    """k nearest neighbours including self -> (knn_ind (n,k) int, knn_dist (n,k) float64), ascending."""
    X = np.asarray(X)
    d = X.shape[1]
    if method is None:
        method = "kdtree" if d <= 5 else "brute"
    if method == "annoy":
        method = "brute"
    if method not in ("kdtree", "brute"):
        raise ValueError("Invalid choice of knnsearch method " + str(method))
    if similarity not in ("angular", "euclidean"):
        raise ValueError("Invalid choice of similarity " + str(similarity))
    if method == "kdtree":
        Y = X / np.linalg.norm(X, axis=1)[:, None] if similarity == "angular" else X
        tree = spatial.cKDTree(Y)
        knn_dist, knn_ind = tree.query(Y, k=k)
    else:
        from . import knn_gpu
        knn_ind, knn_dist = knn_gpu.knnsearch_gpu(X, k, similarity=similarity)
    if dataset is not None:                                       # weightmatrix.py:416-427
        os.makedirs(knn_dir, exist_ok=True)
        np.savez_compressed(os.path.join(knn_dir, dataset.lower() + "_" + metric.lower() + ".npz"), J=knn_ind, D=knn_dist)
    return knn_ind, knn_dist


def load_knn_data(dataset, metric="raw"):
    """Load a stored kNN search (fields J, D).  Reference weightmatrix.py:431-467; nothing is downloaded here."""
    path = os.path.join(knn_dir, dataset.lower() + "_" + metric.lower() + ".npz")
    if not os.path.exists(path):
        raise FileNotFoundError("no stored kNN data at %s (the B200 backend never downloads)" % path)
    M = np.load(path, allow_pickle=True)
    return M["J"], M["D"]


def knn(data, k, kernel="gaussian", eta=None, symmetrize=True, metric="raw", similarity="euclidean", knn_data=None):
    """kNN weight matrix as a scipy CSR (float64, zero diagonal).  Reference weightmatrix.py:68-187."""
    k += 1                                                        # :119 (self is counted)
    if knn_data is not None:
        knn_ind, knn_dist = knn_data
    elif type(data) is str:
        knn_ind, knn_dist = load_knn_data(data, metric=metric)   # :123-124
    else:
        knn_ind, knn_dist = knnsearch(data, k, similarity=similarity)
    knn_ind = np.asarray(knn_ind)
    knn_dist = np.asarray(knn_dist, dtype=np.float64)
    n = knn_ind.shape[0]
    k = np.minimum(knn_ind.shape[1], k)
    knn_ind = knn_ind[:, :k]
    knn_dist = knn_dist[:, :k]
    if eta is None:
        if kernel == "uniform":
            weights = np.ones_like(knn_dist)
        elif kernel == "gaussian":
            D = knn_dist * knn_dist
            eps = D[:, k - 1]
            weights = np.exp(-4 * D / eps[:, None])
        elif kernel == "symgaussian":
            eps = knn_dist[:, k - 1]
            weights = np.exp(-4 * knn_dist * knn_dist / eps[:, None] / eps[knn_ind])
        elif kernel == "distance":
            weights = knn_dist
        elif kernel == "singular":
            weights = knn_dist.copy()
            weights[knn_dist == 0] = 1
            weights = 1 / weights
        else:
            raise ValueError("Invalid choice of kernel: " + str(kernel))
    else:
        D = knn_dist * knn_dist
        eps = D[:, k - 1]
        weights = eta(D / eps)
    if n >= _device_assembly_min_n:
        # COO -> CSR, symmetrisation by the kernel's rule (also when eta is given, as in the reference), zero diagonal:
        # on the device, bit-identical to the scipy expressions below
        rule = 2 if kernel in ["distance", "uniform", "singular"] else 3 if kernel == "symgaussian" else 1
        return _assemble_on_device(knn_ind, weights, n, k, rule if symmetrize else 0)
    knn_ind = knn_ind.flatten()
    weights = weights.flatten()
    self_ind = (np.ones((n, k)) * np.arange(n)[:, None]).flatten()
    W = sparse.coo_matrix((weights, (self_ind, knn_ind)), shape=(n, n)).tocsr()
    if symmetrize:
        if kernel in ["distance", "uniform", "singular"]:
            W = sparse_max(W, W.transpose())
        elif kernel == "symgaussian":
            W = W + W.T.multiply(W.T > W) - W.multiply(W.T > W)
        else:
            W = (W + W.transpose()) / 2
    W = sparse.csr_matrix(W)
    W.setdiag(0)
    W.eliminate_zeros()
    return W
