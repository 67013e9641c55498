"""Host-side helpers on the hot path (mirror of reference graphlearning/utils.py)."""
from __future__ import annotations

import ctypes

import numpy as np
from scipy import sparse


def class_priors(labels):
    """Fraction of the (non-negative) labels in each class.  Reference graphlearning/utils.py:117-142."""
    labels = np.asarray(labels)
    classes = np.unique(labels)
    classes = classes[classes >= 0]
    return np.array([np.sum(labels == cls) for cls in classes]) / np.sum(labels >= 0)


def numpy_load(file, field):
    """One array of an .npz file.  Reference graphlearning/utils.py:219-239 (sys.exit there, an exception here)."""
    try:
        return np.load(file, allow_pickle=True)[field]
    except Exception as e:
        raise OSError("Error: Cannot open " + str(file) + ".") from e


def labels_to_onehot(labels, k=None):
    """One-hot encoding, width max(k, max(label)+1).  Reference graphlearning/utils.py:536-572."""
    labels = np.asarray(labels)
    n = labels.shape[0]
    kk = int(np.max(labels)) + 1
    if k is not None:
        kk = max(kk, int(k))
    labels = labels.astype(int)
    onehot = np.zeros((n, kk))
    onehot[range(n), labels] = 1
    return onehot


def _boundary_handling(bdy_set, bdy_val):
    """Boolean mask / list / scalar boundary data -> index and value arrays.  Reference graphlearning/utils.py:144-173."""
    if type(bdy_set) == list:
        bdy_set = np.array(bdy_set)
    if bdy_set.dtype == bool:
        bdy_set = np.where(bdy_set)[0]
    m = len(bdy_set)
    if type(bdy_val) != np.ndarray:
        bdy_val = np.ones((m,)) * bdy_val
    return bdy_set, bdy_val


def conjgrad(A, b, x0=None, max_iter=1e5, tol=1e-10, return_info=False):
    """Conjugate gradient for A x = b with one or several right-hand sides, on the GPU (cg.cu).

    Mirror of reference graphlearning/utils.py:483-532: same arguments, same stopping rule
    (sqrt of the residual energy summed over ALL columns <= tol, at least one iteration), b may be (n,) or
    (n,c); returns x with the shape of b (float64).  return_info=True also returns (iterations, err, launches).
    """
    from . import _lib
    A = sparse.csr_matrix(A)
    b = np.asarray(b, dtype=np.float64)
    one_d = b.ndim == 1
    B = np.ascontiguousarray(b.reshape(b.shape[0], -1))
    n, c = B.shape
    if A.shape != (n, n):
        raise ValueError("A must be %d x %d" % (n, n))
    rp = np.ascontiguousarray(A.indptr, dtype=np.int32)
    col = np.ascontiguousarray(A.indices, dtype=np.int32)
    val = np.ascontiguousarray(A.data, dtype=np.float64)
    X0 = None if x0 is None else np.ascontiguousarray(np.asarray(x0, dtype=np.float64).reshape(n, c))
    x = np.empty((n, c), dtype=np.float64)
    iters, err, nl = ctypes.c_int64(0), ctypes.c_double(0.0), ctypes.c_int(0)
    as_p = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None
    max_iter = int(min(float(max_iter), 2.0 ** 62))
    _lib.call("glb_cg_host", as_p(rp), as_p(col), as_p(val), n, len(col), as_p(B), as_p(X0), c, float(tol), max_iter,
              as_p(x), ctypes.byref(iters), ctypes.byref(err), ctypes.byref(nl))
    if one_d:
        x = x[:, 0]
    if return_info:
        return x, (iters.value, err.value, nl.value)
    return x


def randomized_svd(A, k=10, c=None, q=1):
    """Randomized SVD on the GPU; mirror of reference graphlearning/utils.py:576-642 (see spectral.randomized_svd)."""
    from . import spectral
    return spectral.randomized_svd(A, k=k, c=c, q=q)
