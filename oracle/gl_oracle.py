"""CPU oracle for the GraphLearning Poisson/Laplace hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``graphlearning_b200/`` may import this
module; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / baseline.

Each function restates, in fp64 numpy/scipy, one function of the reference
(jwcalder/GraphLearning v1.7.5).  File:line citations are relative to the
reference checkout.  The reference itself is pure Python on this path and its
arithmetic happens inside ``scipy.sparse`` (``csr_matvecs`` / ``csr_matvec``) and
``scipy.spatial.cKDTree`` - third-party packages the reference does not pin
(requirements.txt:1-4); the versions in this image are scipy 1.18.1 / numpy 2.3.5
and this oracle calls the same entry points.  An independent plain-C restatement
of the ``csr_matvecs`` row loop and of the Poisson iterate lives in
``oracle/oracle_kernels.c`` and is cross-checked against this file in
``tests/test_oracle.py``.

Pinning: the reference ships no tests or golden vectors, so the
oracle is pinned against outputs of the reference itself, generated in the build
container by ``oracle/make_golden.py`` (imports /root/reference with a matplotlib
stub) and committed under ``tests/golden/``.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse, spatial


# ----------------------------------------------------------------------------
# kNN search  (graphlearning/weightmatrix.py:297-361)
# ----------------------------------------------------------------------------
def knnsearch(X, k, method="kdtree", similarity="euclidean"):
    """Exact kNN incl. self.  weightmatrix.py:339-361 (kdtree :349-352, brute :354-361)."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    if similarity not in ("angular", "euclidean"):
        raise ValueError("Invalid choice of similarity " + similarity)
    if similarity == "angular":                                   # :344-345
        Y = X / np.linalg.norm(X, axis=1)[:, None]
    else:
        Y = X
    if method == "kdtree":                                        # :349-352
        tree = spatial.cKDTree(Y)
        knn_dist, knn_ind = tree.query(Y, k=k)
    elif method == "brute":                                       # :354-361
        knn_ind = np.zeros((n, k), dtype=int)
        knn_dist = np.zeros((n, k))
        for i in range(n):
            dist = np.linalg.norm(Y - Y[i, :], axis=1)
            knn_ind[i, :] = np.argsort(dist)[:k]
            knn_dist[i, :] = dist[knn_ind[i, :]]
    else:
        raise ValueError("Invalid choice of knnsearch method " + method)
    return knn_ind, knn_dist


def knnsearch_rows(X, rows, k):
    """Brute-force kNN of a subset of query rows against all of X (same arithmetic as
    weightmatrix.py:359-361, ties broken by index through a stable sort).  Used for
    full-size spot checks where the whole O(n^2 d) loop is too slow on CPU."""
    X = np.asarray(X, dtype=np.float64)
    rows = np.asarray(rows)
    ind = np.zeros((len(rows), k), dtype=np.int64)
    dst = np.zeros((len(rows), k))
    for t, i in enumerate(rows):
        dist = np.linalg.norm(X - X[i, :], axis=1)
        o = np.argsort(dist, kind="stable")[:k]
        ind[t] = o
        dst[t] = dist[o]
    return ind, dst


# ----------------------------------------------------------------------------
# weight matrix  (graphlearning/weightmatrix.py:68-187, utils.py:263-286)
# ----------------------------------------------------------------------------
def sparse_max(A, B):
    """utils.py:263-286."""
    I = (A + B) > 0
    IB = B > A
    IA = I - IB
    return A.multiply(IA) + B.multiply(IB)


def knn_weights(knn_ind, knn_dist, k, kernel="gaussian", symmetrize=True, eta=None):
    """(ind, dist) -> symmetrised CSR weight matrix.  weightmatrix.py:119-187."""
    k = k + 1                                                     # :119
    knn_ind = np.asarray(knn_ind)
    knn_dist = np.asarray(knn_dist, dtype=np.float64)
    n = knn_ind.shape[0]
    k = min(knn_ind.shape[1], k)                                  # :135
    knn_ind = knn_ind[:, :k]
    knn_dist = knn_dist[:, :k]
    if eta is None:
        if kernel == "uniform":                                   # :142-143
            weights = np.ones_like(knn_dist)
        elif kernel == "gaussian":                                # :144-147
            D = knn_dist * knn_dist
            eps = D[:, k - 1]
            weights = np.exp(-4 * D / eps[:, None])
        elif kernel == "symgaussian":                             # :148-150
            eps = knn_dist[:, k - 1]
            weights = np.exp(-4 * knn_dist * knn_dist / eps[:, None] / eps[knn_ind])
        elif kernel == "distance":                                # :151-152
            weights = knn_dist
        elif kernel == "singular":                                # :153-156
            weights = knn_dist.copy()
            weights[knn_dist == 0] = 1
            weights = 1 / weights
        else:
            raise ValueError("Invalid choice of kernel: " + kernel)
    else:                                                         # :161-164
        D = knn_dist * knn_dist
        eps = D[:, k - 1]
        weights = eta(D / eps)
    knn_ind = knn_ind.flatten()
    weights = weights.flatten()
    self_ind = (np.ones((n, k)) * np.arange(n)[:, None]).flatten()  # :171-172
    W = sparse.coo_matrix((weights, (self_ind, knn_ind)), shape=(n, n)).tocsr()  # :175
    if symmetrize:                                                # :177-183
        if kernel in ["distance", "uniform", "singular"]:
            W = sparse_max(W, W.transpose())
        elif kernel == "symgaussian":
            W = W + W.T.multiply(W.T > W) - W.multiply(W.T > W)
        else:
            W = (W + W.transpose()) / 2
    W = sparse.csr_matrix(W)
    W.setdiag(0)                                                  # :185
    W.eliminate_zeros()                                           # :186
    return W


def knn(X, k, kernel="gaussian", symmetrize=True, similarity="euclidean", method="kdtree"):
    """weightmatrix.knn with an exact search (the reference default for d>5 is the
    approximate, absent ``annoy``)."""
    ind, dist = knnsearch(X, k + 1, method=method, similarity=similarity)
    return knn_weights(ind, dist, k, kernel=kernel, symmetrize=symmetrize)


# ----------------------------------------------------------------------------
# graph normalisation  (graphlearning/graph.py:108-122, 210-233, 469-513)
# ----------------------------------------------------------------------------
def degree_vector(W):
    """graph.py:121."""
    return W * np.ones(W.shape[0])


def degree_matrix(W, p=1):
    """graph.py:229-233."""
    n = W.shape[0]
    d = degree_vector(W)
    return sparse.spdiags(d ** p, 0, n, n).tocsr()


def laplacian(W, normalization="combinatorial"):
    """graph.py:499-513 (coifmanlafon omitted: not on the path)."""
    n = W.shape[0]
    I = sparse.identity(n)
    D = degree_matrix(W)
    if normalization == "combinatorial":
        L = D - W
    elif normalization == "randomwalk":
        L = I - degree_matrix(W, -1) * W
    elif normalization == "normalized":
        Dinv2 = degree_matrix(W, -0.5)
        L = I - Dinv2 * W * Dinv2
    else:
        raise ValueError("Invalid option for graph Laplacian normalization.")
    return L.tocsr()


def ccode_triplets(W):
    """Row-sorted COO triplets the reference hands to its C code.  graph.py:69-84."""
    W = sparse.csr_matrix(W)
    I, J, V = sparse.find(W)
    ind = np.argsort(I, kind="stable")
    I, J, V = I[ind], J[ind], V[ind]
    return (np.ascontiguousarray(I, dtype=np.int32), np.ascontiguousarray(J, dtype=np.int32),
            np.ascontiguousarray(V, dtype=np.float64))


# ----------------------------------------------------------------------------
# labels / predict  (graphlearning/utils.py:536-572, ssl.py:230-266, 1795-1834)
# ----------------------------------------------------------------------------
def labels_to_onehot(labels, k):
    """utils.py:557-570."""
    labels = np.asarray(labels)
    n = labels.shape[0]
    k = max(int(np.max(labels)) + 1, k)
    labels = labels.astype(int)
    onehot = np.zeros((n, k))
    onehot[range(n), labels] = 1
    return onehot


def predict(prob, similarity=True, w=1):
    """ssl.py:257-266."""
    scores = prob - np.min(prob)
    scores = scores / np.max(scores)
    if similarity:
        return np.argmax(scores * w, axis=1)
    return np.argmin(scores * w, axis=1)


def ssl_accuracy(pred_labels, true_labels, train_ind):
    """ssl.py:1819-1834."""
    mask = np.ones(len(pred_labels), dtype=bool)
    mask[train_ind] = False
    p = pred_labels[mask]
    t = true_labels[mask]
    I = t >= 0
    return 100 * np.mean(p[I] == t[I])


# ----------------------------------------------------------------------------
# conjugate gradient  (graphlearning/utils.py:483-532)
# ----------------------------------------------------------------------------
def conjgrad(A, b, x0=None, max_iter=1e5, tol=1e-10, return_iters=False):
    """Multi-RHS CG with per-column alpha/beta and ONE global stopping norm.  utils.py:510-532."""
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - A @ x
    p = r.copy()
    rsold = np.sum(r ** 2, axis=0)
    err = 1
    i = 0
    while (err > tol) and (i < max_iter):
        i += 1
        Ap = A @ p
        alpha = rsold / np.sum(p * Ap, axis=0)
        x += alpha * p
        r -= alpha * Ap
        rsnew = np.sum(r ** 2, axis=0)
        err = np.sqrt(np.sum(rsnew))
        p = r + (rsnew / rsold) * p
        rsold = rsnew
    if return_iters:
        return x, i
    return x


# ----------------------------------------------------------------------------
# Poisson learning  (graphlearning/ssl.py:608-677)
# ----------------------------------------------------------------------------
def poisson_source(n, train_ind, train_labels):
    """ssl.py:611-622."""
    k = len(np.unique(train_labels))
    onehot = labels_to_onehot(train_labels, k)
    source = np.zeros((n, onehot.shape[1]))
    source[train_ind] = onehot - np.mean(onehot, axis=0)
    return source, k


def poisson_gd_setup(W, train_ind, train_labels):
    """Everything before the loop of the gradient-descent branch.  ssl.py:610-645."""
    W = sparse.csr_matrix(W)
    n = W.shape[0]
    W = W - sparse.spdiags(W.diagonal(), 0, n, n)                 # :616
    W = sparse.csr_matrix(W)
    source, k = poisson_source(n, train_ind, train_labels)
    D = degree_matrix(W, p=-1)                                    # :634
    P = D * W.transpose()                                         # :635
    Db = D * source                                               # :636
    v = np.zeros(n)
    v[train_ind] = 1
    v = v / np.sum(v)                                             # :639-641
    deg = degree_vector(W)
    vinf = deg / np.sum(deg)                                      # :642-643
    RW = W.transpose() * D                                        # :644
    return dict(n=n, k=k, P=sparse.csr_matrix(P), Db=Db, v=v, vinf=vinf, RW=sparse.csr_matrix(RW),
                width=source.shape[1])


def poisson_gd(W, train_ind, train_labels, min_iter=50, max_iter=1000, return_iters=False):
    """Poisson learning, solver='gradient_descent', CPU branch.  ssl.py:631-670."""
    s = poisson_gd_setup(W, train_ind, train_labels)
    n, P, Db, v, vinf, RW = s["n"], s["P"], s["Db"], s["v"], s["vinf"], s["RW"]
    u = np.zeros((n, s["k"]))                                     # :645 (k columns; Db may be wider)
    if u.shape[1] != Db.shape[1]:
        u = np.zeros_like(Db)
    T = 0
    while (T < min_iter or np.max(np.absolute(v - vinf)) > 1 / n) and (T < max_iter):   # :667
        u = Db + P * u                                            # :668
        v = RW * v                                                # :669
        T = T + 1
    if return_iters:
        return u, T
    return u


def torch_sparse(A):
    """scipy sparse -> torch COO float32 (utils.py:288-317; torch.sparse.FloatTensor(i, v, size) in today's spelling)."""
    import torch
    A = A.tocoo()
    i = torch.from_numpy(np.vstack((A.row, A.col)).astype(np.int64))
    v = torch.from_numpy(A.data.astype(np.float32))
    return torch.sparse_coo_tensor(i, v, torch.Size(A.shape))


def poisson_gd_use_cuda(W, train_ind, train_labels, min_iter=50, max_iter=1000, return_iters=False, host_mixing=True):
    """The reference's OWN GPU variant of the gradient-descent branch, `use_cuda=True` (ssl.py:649-663): P as a torch COO
    fp32 matrix, u <- torch.sparse.addmm(Db, P, u) on the device while the stopping vector v <- RW*v and its np.max stay on
    the host in every iteration.  Needs a CUDA device; used by bench.py as the "GPU baseline to beat" (SURVEY 8a row a14).
    host_mixing=False leaves the host-side v update out (an upper bound of what the torch path could do)."""
    import torch
    s = poisson_gd_setup(W, train_ind, train_labels)
    n, P, Db, v, vinf, RW = s["n"], s["P"], s["Db"], s["v"], s["vinf"], s["RW"]
    Pt = torch_sparse(P).cuda()                                   # :653
    ut = torch.from_numpy(np.zeros_like(Db)).float().cuda()       # :654
    Dbt = torch.from_numpy(Db).float().cuda()                     # :655
    T = 0
    if host_mixing:
        while (T < min_iter or np.max(np.absolute(v - vinf)) > 1 / n) and (T < max_iter):   # :657
            ut = torch.sparse.addmm(Dbt, Pt, ut)                  # :658
            v = RW * v                                            # :659
            T = T + 1
    else:
        while T < max_iter:
            ut = torch.sparse.addmm(Dbt, Pt, ut)
            T = T + 1
    u = ut.cpu().numpy()                                          # :663
    if return_iters:
        return u, T
    return u


def poisson_cg(W, train_ind, train_labels, tol=1e-3, return_iters=False):
    """Poisson learning, default solver='conjugate_gradient'.  ssl.py:624-629."""
    W = sparse.csr_matrix(W)
    n = W.shape[0]
    W = sparse.csr_matrix(W - sparse.spdiags(W.diagonal(), 0, n, n))
    source, _ = poisson_source(n, train_ind, train_labels)
    L = laplacian(W, "normalized")
    D = degree_matrix(W, p=-0.5)
    x, it = conjgrad(L, D * source, tol=tol, return_iters=True)
    u = D * x
    if return_iters:
        return u, it
    return u


# ----------------------------------------------------------------------------
# Laplace learning  (graphlearning/ssl.py:1206-1261)
# ----------------------------------------------------------------------------
def laplace_system(W, train_ind, train_labels, normalization="combinatorial", tau=0.0):
    """Assemble the Jacobi-scaled Dirichlet system.  ssl.py:1217-1246 (order=1, no reweighting)."""
    W = sparse.csr_matrix(W)
    n = W.shape[0]
    k = len(np.unique(train_labels))
    tau_v = np.ones(n) * tau if np.isscalar(tau) else np.asarray(tau, dtype=np.float64)
    L = sparse.spdiags(tau_v, 0, n, n) + laplacian(W, normalization)      # :1222
    L = sparse.csr_matrix(L)
    F = labels_to_onehot(train_labels, k)                         # :1229
    idx = np.full((n,), True, dtype=bool)
    idx[train_ind] = False                                        # :1232-1233
    b = -L[:, train_ind] * F                                      # :1236
    b = b[idx, :]
    A = L[idx, :]
    A = A[:, idx]                                                 # :1240-1241
    m = A.shape[0]
    Mdiag = 1 / np.sqrt(A.diagonal() + 1e-10)                     # :1245-1246
    M = sparse.spdiags(Mdiag, 0, m, m).tocsr()
    return dict(n=n, k=k, F=F, idx=idx, A=A, b=b, M=M, MAM=sparse.csr_matrix(M * A * M), Mb=M * b)


def laplace_fit(W, train_ind, train_labels, normalization="combinatorial", tau=0.0, tol=1e-5,
                mean_shift=False, return_iters=False):
    """ssl.py:1206-1261 (reweighting='none', order=1)."""
    s = laplace_system(W, train_ind, train_labels, normalization, tau)
    v, it = conjgrad(s["MAM"], s["Mb"], tol=tol, return_iters=True)       # :1249
    v = s["M"] * v                                                # :1250
    u = np.zeros((s["n"], s["k"])) if s["F"].shape[1] == s["k"] else np.zeros((s["n"], s["F"].shape[1]))
    u[s["idx"], :] = v
    u[train_ind, :] = s["F"]                                      # :1253-1255
    if mean_shift:
        u -= np.mean(u, axis=0)
    if return_iters:
        return u, it
    return u


# ----------------------------------------------------------------------------
# p-Laplace iterates  (c_code/lp_iterate.cpp:35-187) - slow pure-Python restatement,
# small cases only; the compiled restatement is oracle_kernels.c, the compiled
# reference is oracle/_ref/liblp_ref.so.
# ----------------------------------------------------------------------------
def _row_starts(J, n, M):
    """lp_iterate.cpp:44-57 (start/num scan over row-sorted J)."""
    start = np.zeros(n, dtype=np.int64)
    num = np.zeros(n, dtype=np.int64)
    j = 0
    for i in range(n):
        start[i] = j
        while j < M and J[j] == i:
            num[i] += 1
            j += 1
    return start, num


def lp_iterate(uu, ul, I, J, W, ind, val, p, T, tol):
    """Jacobi p-Laplace barrier iteration.  c_code/lp_iterate.cpp:35-125.
    Mutates nothing; returns (uu, ul) as the caller's buffers would hold them on return
    (the C code swaps pointers each sweep, Appendix A.9 of SURVEY.md), plus sweeps run."""
    n, M, m = len(uu), len(I), len(ind)
    alpha = 1 / p
    delta = 1 - 2 / p
    dt = 0.9 / (alpha + 2 * delta)
    start, num = _row_starts(J, n, M)
    invdeg = np.zeros(n)
    for i in range(n):
        invdeg[i] = alpha / np.sum(W[start[i]:start[i] + num[i]]) if num[i] > 0 else np.inf
    dt = dt / np.max(W)
    bufs = {"uu": uu.copy(), "ul": ul.copy(), "vu": np.zeros(n), "vl": np.zeros(n)}
    a_u, a_l, b_u, b_l = "uu", "ul", "vu", "vl"
    sweeps = 0
    for it in range(T):
        sweeps += 1
        cu, cl = bufs[a_u], bufs[a_l]
        nu, nl = bufs[b_u], bufs[b_l]
        err = 0.0
        for i in range(n):
            s, e = start[i], start[i] + num[i]
            du = W[s:e] * (cu[I[s:e]] - cu[i])
            nu[i] = cu[i] + dt * (invdeg[i] * np.sum(du) + delta * (min(du.min(initial=0), 0) + max(du.max(initial=0), 0)))
            dl = W[s:e] * (cl[I[s:e]] - cl[i])
            nl[i] = cl[i] + dt * (invdeg[i] * np.sum(dl) + delta * (min(dl.min(initial=0), 0) + max(dl.max(initial=0), 0)))
            err = max(cu[i] - cl[i], err)
        nu[ind] = val
        nl[ind] = val
        if err < tol and it > 10:
            break
        a_u, b_u = b_u, a_u
        a_l, b_l = b_l, a_l
    return bufs["uu"], bufs["ul"], sweeps


def lip_iterate(u, I, J, W, ind, val, T, tol, alpha, beta):
    """Gauss-Seidel game-theoretic p-Laplace sweeps.  c_code/lp_iterate.cpp:129-187."""
    u = u.copy()
    n, M = len(u), len(I)
    start, num = _row_starts(J, n, M)
    mask = np.ones(n, dtype=bool)
    u[ind] = val
    mask[ind] = False
    sweeps = 0
    for it in range(T):
        sweeps += 1
        err = 0.0
        for i in range(n):
            if mask[i]:
                s, e = start[i], start[i] + num[i]
                nb = u[I[s:e]]
                ne = alpha * np.sum(W[s:e] * nb) / np.sum(W[s:e]) + beta * (nb.min() + nb.max()) / 2
                err = max(abs(u[i] - ne), err)
                u[i] = ne
        if err < tol and it > 20:
            break
    return u, sweeps


# ---------------------------------------------------------------------------------------------------
# spectral decomposition, reweighting  (rows added with the spectral / p-Laplace kernels)
# ---------------------------------------------------------------------------------------------------
def eigen_decomp(W, normalization="combinatorial", k=10, tol=0):
    """graph.eigen_decomp, method='exact', gamma=0.  graphlearning/graph.py:728-765: ARPACK svds of the shifted /
    normalised matrix, vals = shift - s sorted ascending, random-walk vectors rescaled by D^-1/2."""
    from scipy.sparse import linalg as splinalg
    W = sparse.csr_matrix(W)
    n = W.shape[0]
    if normalization in ("randomwalk", "normalized"):
        D = degree_matrix(W, p=-0.5)
        A = D * W * D
        u, s, vt = splinalg.svds(A, k=k, tol=tol)
        vals = 1 - s
        ind = np.argsort(vals)
        vals, vecs = vals[ind], u[:, ind]
        if normalization == "randomwalk":
            vecs = D @ vecs
        return vals, vecs
    if normalization == "combinatorial":
        L = laplacian(W)
        M = 2 * np.max(degree_vector(W))
        A = M * sparse.identity(n) - L
        u, s, vt = splinalg.svds(A, k=k, tol=tol)
        vals = M - s
        ind = np.argsort(vals)
        return vals[ind], u[:, ind]
    raise ValueError("Invalid choice of normalization")


def poisson_spectral(W, train_ind, train_labels, spectral_cutoff=10, p=1):
    """ssl.poisson._fit, solver='spectral'.  graphlearning/ssl.py:615-617, 680-688."""
    W = sparse.csr_matrix(W)
    n = W.shape[0]
    source, k = poisson_source(n, train_ind, train_labels)
    W0 = sparse.csr_matrix(W - sparse.spdiags(W.diagonal(), 0, n, n))
    vals, vecs = eigen_decomp(W0, normalization="randomwalk", k=spectral_cutoff + 1)
    V, vals = vecs[:, 1:], vals[1:]
    if p != 1:
        vals = vals ** p
    L = sparse.spdiags(1 / vals, 0, spectral_cutoff, spectral_cutoff)
    return V @ (L @ (V.T @ source))


def reweight(W, idx, method="poisson", normalization="combinatorial", X=None, alpha=2, zeta=1e7, r=0.1):
    """graph.reweight.  graphlearning/graph.py:412-462."""
    from scipy import spatial
    W = sparse.csr_matrix(W)
    n = W.shape[0]
    if method == "poisson":
        f = np.zeros(n)
        f[idx] = 1
        if normalization == "combinatorial":
            f -= np.mean(f)
            L = laplacian(W)
        else:
            d = degree_vector(W) ** 0.5
            f -= np.sum(d * f) / np.sum(d)
            L = laplacian(W, normalization=normalization)
        w = conjgrad(L, f, tol=1e-5)
        w -= np.min(w)
        w += 1e-5
        D = sparse.spdiags(w, 0, n, n).tocsr()
        return D * W * D
    if method == "wnll":
        a = np.ones((n,))
        a[idx] = n / len(idx)
        D = sparse.spdiags(a, 0, n, n).tocsr()
        return D * W + W * D
    if method == "properly":
        rzeta = r / (zeta - 1) ** (1 / alpha)
        Dn, _ = spatial.cKDTree(X[idx, :]).query(X)
        Dn[Dn < rzeta] = rzeta
        D = sparse.spdiags(1 + (r / Dn) ** alpha, 0, n, n).tocsr()
        return D * W + W * D
    raise ValueError("Invalid reweighting method")


# ----------------------------------------------------------------------------
# synthetic workloads - shared by tests and bench so that both arms
# of every comparison see identical inputs.
# ----------------------------------------------------------------------------
def synthetic_blobs(n, d, c=10, seed=0, dtype=np.float32):
    """Config 2/3 generator: c centres ~N(0, 3^2 I_d), X = centre[label] + N(0, I_d)."""
    rng = np.random.default_rng(seed)
    centres = rng.normal(0.0, 3.0, size=(c, d))
    labels = rng.integers(0, c, n)
    X = centres[labels] + rng.normal(0.0, 1.0, size=(n, d))
    return X.astype(dtype), labels.astype(np.int64)


def one_per_class(labels, rate=1, seed=0):
    """Deterministic stand-in for trainsets.generate(labels, rate) (trainsets.py:121-131)."""
    rng = np.random.default_rng(seed)
    out = []
    for l in np.unique(labels):
        out += rng.choice(np.flatnonzero(labels == l), size=rate, replace=False).tolist()
    return np.array(out, dtype=np.int64)
