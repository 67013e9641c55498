"""Graph object (mirror of the hot-path part of reference graphlearning/graph.py:25-84,108-122,210-233,469-513).

Holds the weight matrix as a scipy CSR on the host - the same object the reference hands around - and
builds the degree / Laplacian matrices with the same scipy expressions; the arithmetic that matters for
the iterate (degrees, D^-1 scaling, transposition) is redone on the device from the CSR when a solver
runs (graph_ops.cu), so nothing here is on the timed path.
"""
from __future__ import annotations

import ctypes

import numpy as np
from scipy import sparse


class graph:
    def __init__(self, W, labels=None, features=None, label_names=None, node_names=None):
        self.weight_matrix = sparse.csr_matrix(W)
        self.labels = labels
        self.features = features
        self.num_nodes = W.shape[0]
        self.label_names = label_names
        self.node_names = node_names
        self._device = {}
        self._coo = None
        self.eigendata = {}
        for norm in ("normalized", "randomwalk", "combinatorial"):
            self.eigendata[norm] = dict(eigenvectors=None, eigenvalues=None, method=None, k=None, c=None, gamma=None,
                                        tol=None, q=None)

    def _ccode_arrays(self):
        """Row-sorted COO triplets (I=row, J=column, V=weight) built with the reference's own expressions
        (graph.__ccode_init__, graph.py:69-84) - the stored order inside a row is what the sweeps sum in.
        Built on first use instead of on every construction (0.46 s at n = 70k in the reference)."""
        W = self.weight_matrix
        key = (id(W), W.nnz, id(W.data))
        if self._coo is None or self._coo[0] != key:
            if W.has_canonical_format:
                # sparse.find of a canonical CSR lists the stored nonzeros row by row with ascending columns, and the
                # argsort of that already sorted row index is the identity: same triplets without the 0.4 s of COO work
                keep = W.data != 0
                I = np.repeat(np.arange(W.shape[0]), np.diff(W.indptr))[keep]
                J, V = W.indices[keep], W.data[keep]
            else:
                I, J, V = sparse.find(W)
                ind = np.argsort(I)
                I, J, V = I[ind], J[ind], V[ind]
            self._coo = (key, np.ascontiguousarray(I, dtype=np.int32), np.ascontiguousarray(J, dtype=np.int32),
                         np.ascontiguousarray(V, dtype=np.float64))
        return self._coo[1:]

    @property
    def I(self):
        return self._ccode_arrays()[0]

    @property
    def J(self):
        return self._ccode_arrays()[1]

    @property
    def V(self):
        return self._ccode_arrays()[2]

    def poisson_handle(self):
        """Device-resident Poisson state of this graph's weight matrix (built on first use, then shared by every
        model/fit on the graph; rebuilt if weight_matrix is replaced)."""
        from . import device
        W = self.weight_matrix
        key = (id(W), W.nnz, id(W.data))
        ent = self._device.get("poisson")
        if ent is None or ent[0] != key:
            ent = (key, device.PoissonGraphHandle(W))
            self._device["poisson"] = ent
        return ent[1]

    def laplace_handle(self, normalization, tau):
        """Device-resident W + Laplacian scalings for ssl.laplace fits on this graph (one per normalisation and tau;
        rebuilt if weight_matrix is replaced)."""
        from . import device
        W = self.weight_matrix
        tau_key = None if tau is None else tau.tobytes()
        key = (id(W), W.nnz, id(W.data), normalization, tau_key)
        ent = self._device.get("laplace")
        if ent is None or ent[0] != key:
            left, right, diag = self._laplacian_scalings(normalization)
            Wc, rp, ci, val = self._canonical_weights()
            ent = (key, device.LaplaceGraphHandle(rp, ci, val, self.num_nodes, left, right, diag, tau))
            self._device["laplace"] = ent
        return ent[1]

    def degree_vector(self):
        """graph.py:108-122."""
        return self.weight_matrix * np.ones(self.num_nodes)

    def degree_matrix(self, p=1):
        """graph.py:210-233."""
        n = self.num_nodes
        d = self.degree_vector()
        return sparse.spdiags(d ** p, 0, n, n).tocsr()

    def adjacency(self):
        A = self.weight_matrix.copy()
        A.data[:] = 1
        return A

    def _laplacian_scalings(self, normalization):
        """(left, right, diag) of L = Diag(diag) - Diag(left) W Diag(right), with d^p computed as the reference computes
        it (numpy power of the degree vector, graph.py:230)."""
        d = self.degree_vector()
        if normalization == "combinatorial":
            return None, None, d
        one = np.ones(self.num_nodes)
        if normalization == "randomwalk":
            return d ** -1, None, one
        if normalization == "normalized":
            dl = d ** -0.5
            return dl, dl, one
        raise ValueError("Invalid option for graph Laplacian normalization.")

    def _canonical_weights(self):
        W = self.weight_matrix
        if not W.has_canonical_format:
            W = W.copy()
            W.sum_duplicates()
        return (W, np.ascontiguousarray(W.indptr, dtype=np.int32), np.ascontiguousarray(W.indices, dtype=np.int32),
                np.ascontiguousarray(W.data, dtype=np.float64))

    def laplacian(self, normalization="combinatorial", alpha=1):
        """graph.py:469-513.  The O(nnz) assembly runs on the device (glb_laplacian_csr_host, csrc/laplace.cu): same
        values as the reference's scipy expressions bit for bit, canonical CSR."""
        from . import _lib
        if normalization == "coifmanlafon":
            D = self.degree_matrix(p=-alpha)
            return graph(D * self.weight_matrix * D).laplacian(normalization="randomwalk")
        left, right, diag = self._laplacian_scalings(normalization)
        W, rp, ci, val = self._canonical_weights()
        n = self.num_nodes
        out_rp = np.empty(n + 1, dtype=np.int32)
        out_ci = np.empty(W.nnz + n, dtype=np.int32)
        out_val = np.empty(W.nnz + n, dtype=np.float64)
        vp = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None
        _lib.call("glb_laplacian_csr_host", vp(rp), vp(ci), vp(val), n, W.nnz, vp(left), vp(right), vp(diag), vp(out_rp),
                  vp(out_ci), vp(out_val))
        m = int(out_rp[n])
        return sparse.csr_matrix((out_val[:m], out_ci[:m], out_rp), shape=(n, n))

    def reweight(self, idx, method="poisson", normalization="combinatorial", tau=0, X=None, alpha=2, zeta=1e7, r=0.1):
        """Reweight the weight matrix near the labelled nodes `idx`.  Reference graphlearning/graph.py:368-466; the
        'poisson' method's linear solve (utils.conjgrad(L, f, tol=1e-5), :430) runs on the GPU CG."""
        from . import utils
        from scipy import spatial
        n = self.num_nodes
        if method == "poisson":
            f = np.zeros(n)
            f[idx] = 1
            if normalization == "combinatorial":
                f -= np.mean(f)
                L = self.laplacian()
            elif normalization == "normalized":
                d = self.degree_vector() ** (0.5)
                c = np.sum(d * f) / np.sum(d)
                f -= c
                L = self.laplacian(normalization=normalization)
            else:
                raise ValueError("Unsupported normalization " + str(normalization) + " for graph.reweight.")
            w = utils.conjgrad(L, f, tol=1e-5)
            w -= np.min(w)
            w += 1e-5
            D = sparse.spdiags(w, 0, n, n).tocsr()
            return D * self.weight_matrix * D
        if method == "wnll":
            m = len(idx)
            a = np.ones((n,))
            a[idx] = n / m
            D = sparse.spdiags(a, 0, n, n).tocsr()
            return D * self.weight_matrix + self.weight_matrix * D
        if method == "properly":
            if X is None:
                raise ValueError("Must provide data features X for properly weighted graph Laplacian.")
            rzeta = r / (zeta - 1) ** (1 / alpha)
            Dn, _ = spatial.cKDTree(X[idx, :]).query(X)
            Dn[Dn < rzeta] = rzeta
            gamma = 1 + (r / Dn) ** alpha
            D = sparse.spdiags(gamma, 0, n, n).tocsr()
            return D * self.weight_matrix + self.weight_matrix * D
        raise ValueError("Invalid reweighting method " + str(method) + ".")

    def _lip_multi(self, bdy_set, bdy_val, T, tol, weighted, alpha, beta):
        """c right-hand sides (bdy_val: m x c) through glb_lip_iterate_multi_host: the classes of a one-vs-rest fit share the
        graph and the Dirichlet rows, so they run as ONE batched sweep kernel; column k equals the single call on column k."""
        from . import _lib
        import ctypes
        n = self.num_nodes
        I, J, V = self._ccode_arrays()
        bs = np.ascontiguousarray(bdy_set, dtype=np.int32)
        bv = np.ascontiguousarray(bdy_val, dtype=np.float64)
        c = bv.shape[1]
        u = np.zeros((n, c), dtype=np.float64)
        sw = (ctypes.c_int * c)()
        nl = ctypes.c_int(0)
        ptr = lambda a: ctypes.c_void_p(a.ctypes.data)
        _lib.call("glb_lip_iterate_multi_host", ptr(u), ptr(J), ptr(I), ptr(V), ptr(bs), ptr(bv), int(T), float(tol),
                  1 if weighted else 0, float(alpha), float(beta), n, len(I), len(bs), c, sw, ctypes.byref(nl))
        self.sweeps, self.gpu_launches = list(sw), nl.value
        return u

    # ---- p-Laplace / AMLE sweeps on the GPU (plaplace.cu) ---------------------------------------------------
    def plaplace(self, bdy_set, bdy_val, p, tol=1e-1, max_num_it=1e6, prog=False, fast=True):
        """Game-theoretic p-Laplace equation with Dirichlet data.  Reference graphlearning/graph.py:1177-1279;
        the sweeps of cextensions.lip_iterate / lp_iterate (c_code/lp_iterate.cpp) run on the GPU, bit-identical.
        Attributes `sweeps` and `gpu_launches` describe the last solve."""
        from . import utils, _lib
        import ctypes
        n = self.num_nodes
        alpha = 1 / (p - 1)
        beta = 1 - alpha
        bdy_set, bdy_val = utils._boundary_handling(bdy_set, bdy_val)
        I, J, V = self._ccode_arrays()
        ptr = lambda a: ctypes.c_void_p(a.ctypes.data)
        sw, nl = ctypes.c_int(0), ctypes.c_int(0)
        T = int(min(float(max_num_it), 2.0 ** 31 - 1))
        if fast and np.ndim(bdy_val) == 2:                       # batched one-vs-rest classes (not in the reference API)
            return self._lip_multi(bdy_set, bdy_val, T, 1e-6, False, alpha, beta)
        if fast:
            u = np.zeros((n,), dtype=np.float64)
            bs = np.ascontiguousarray(bdy_set, dtype=np.int32)
            bv = np.ascontiguousarray(bdy_val, dtype=np.float64)
            tol = 1e-6                                            # graph.py:1259
            _lib.call("glb_lip_iterate_host", ptr(u), ptr(J), ptr(I), ptr(V), ptr(bs), ptr(bv), T, float(tol), 0,
                      float(alpha), float(beta), n, len(I), len(bs), ctypes.byref(sw), ctypes.byref(nl))
        else:
            uu = np.max(bdy_val) * np.ones((n,))
            ul = np.min(bdy_val) * np.ones((n,))
            uu[bdy_set] = bdy_val
            ul[bdy_set] = bdy_val
            uu = np.ascontiguousarray(uu, dtype=np.float64)
            ul = np.ascontiguousarray(ul, dtype=np.float64)
            bs = np.ascontiguousarray(bdy_set, dtype=np.int32)
            bv = np.ascontiguousarray(bdy_val, dtype=np.float64)
            _lib.call("glb_lp_iterate_host", ptr(uu), ptr(ul), ptr(J), ptr(I), ptr(V), ptr(bs), ptr(bv), float(p), T,
                      float(tol), n, len(I), len(bs), ctypes.byref(sw), ctypes.byref(nl))
            u = (uu + ul) / 2
        self.sweeps, self.gpu_launches = sw.value, nl.value
        return u

    def amle(self, bdy_set, bdy_val, tol=1e-5, max_num_it=1000, weighted=True, prog=False):
        """Absolutely minimal Lipschitz extension.  Reference graphlearning/graph.py:1281-1332."""
        from . import utils, _lib
        import ctypes
        n = self.num_nodes
        u = np.zeros((n,), dtype=np.float64)
        bdy_set, bdy_val = utils._boundary_handling(bdy_set, bdy_val)
        if np.ndim(bdy_val) == 2:                                 # batched one-vs-rest classes (not in the reference API)
            return self._lip_multi(bdy_set, bdy_val, int(min(float(max_num_it), 2.0 ** 31 - 1)), tol, weighted, 0.0, 1.0)
        bs = np.ascontiguousarray(bdy_set, dtype=np.int32)
        bv = np.ascontiguousarray(bdy_val, dtype=np.float64)
        I, J, V = self._ccode_arrays()
        ptr = lambda a: ctypes.c_void_p(a.ctypes.data)
        sw, nl = ctypes.c_int(0), ctypes.c_int(0)
        T = int(min(float(max_num_it), 2.0 ** 31 - 1))
        _lib.call("glb_lip_iterate_host", ptr(u), ptr(J), ptr(I), ptr(V), ptr(bs), ptr(bv), T, float(tol),
                  1 if weighted else 0, 0.0, 1.0, n, len(I), len(bs), ctypes.byref(sw), ctypes.byref(nl))
        self.sweeps, self.gpu_launches = sw.value, nl.value
        return u

    # ---- spectral decomposition on the GPU (spectral.cu / spectral.py) --------------------------------------
    def page_rank(self, alpha=0.85, v=None, tol=1e-10):
        """PageRank by the power iteration u <- alpha P u + (1 - alpha) v, P = W^T D^-1, until max |u_new - u| <= tol.
        Reference graphlearning/graph.py:1374-1412; the products run on the device (fp64 block SpMM of spectral.cu, one
        launch per iteration plus the max-norm reduction of mbo.cu), same iteration count as the reference."""
        from . import _lib, device, spectral
        n = self.num_nodes
        u0 = np.ones((n,)) / n
        v = u0.copy() if v is None else np.asarray(v, dtype=np.float64)
        D = self.degree_matrix(p=-1)
        ops = spectral.BlockOps(sparse.csr_matrix(self.weight_matrix.T @ D))
        u, vt = ops.upload(u0[:, None]), ops.upload(v[:, None])
        w = ops.new(1)
        err, it = tol + 1, 0
        e = ctypes.c_double(0.0)
        while err > tol:
            ops.spmm(u, 1, out=w, alpha=alpha, Y2=vt, gamma=1 - alpha)
            _lib.call("glb_max_abs_diff_f64", device.ptr(w), device.ptr(u), n, 1, int(w.shape[1]), int(u.shape[1]), ctypes.byref(e),
                      device.cur_stream())
            err = e.value
            u, w = w, u
            it += 1
        self.page_rank_iterations = it
        self.gpu_launches = ops.launches + it
        return u[:, 0].cpu().numpy()

    def eigen_decomp(self, normalization="combinatorial", method="exact", k=10, c=None, gamma=0, tol=0, q=1):
        """Smallest k eigenpairs of the graph Laplacian.  Reference graphlearning/graph.py:623-806: same shifted /
        normalised matrices, same post-processing (vals = 1 - s or M - s, randomwalk vectors scaled by D^-1/2), same
        result cache.  method='exact' (ARPACK svds in the reference) runs the block Chebyshev subspace iteration,
        method='lowrank' the randomized SVD, both on the device.  The modularity variant (gamma != 0, eigsh on a
        LinearOperator, :772-799) is outside the hot path."""
        from . import spectral
        if c is None:
            c = 2 * k
        ed = self.eigendata[normalization] if normalization in self.eigendata else None
        if ed is None:
            raise ValueError("Invalid choice of normalization")
        if (ed["method"] == method and ed["k"] == k and ed["c"] == c and ed["gamma"] == gamma and ed["tol"] == tol
                and ed["q"] == q):
            return ed["eigenvalues"], ed["eigenvectors"]
        if gamma != 0:
            raise NotImplementedError("eigen_decomp(gamma != 0) (modularity) is not on the B200 hot path")
        if method not in ("exact", "lowrank"):
            raise ValueError("Invalid eigensolver method " + str(method))
        n = self.num_nodes
        if normalization in ("randomwalk", "normalized"):
            D = self.degree_matrix(p=-0.5)
            A = D * self.weight_matrix * D
            shift = 1.0
        else:
            L = self.laplacian()
            shift = 2 * np.max(self.degree_vector())
            A = shift * sparse.identity(n) - L
        if method == "exact":
            u, s, info = spectral.svd_topk(A, k, tol=tol, return_info=True)
        else:
            u, s, vt, info = spectral.randomized_svd(A, k=k, c=c, q=q, return_info=True)
        self.gpu_launches = info["launches"]
        self.eigen_info = info
        if not info.get("converged", True):
            msg = ("eigen_decomp: the block solver stopped at a residual of %.2e (wanted %.2e) after %d outer iterations"
                   % (info["residual"], info["rtol"], info["outer"] + 1))
            if not info.get("stagnated", False):
                # ARPACK raises ArpackNoConvergence at this point (the reference's svds call, graph.py:734); nothing is cached
                raise RuntimeError(msg)
            import warnings
            warnings.warn(msg + " (stagnation at the rounding floor, residual below 1e-9)", RuntimeWarning)
        vals = shift - s
        ind = np.argsort(vals)
        vals = vals[ind]
        vecs = u[:, ind]
        if normalization == "randomwalk":
            vecs = D @ vecs
        ed.update(method=method, k=k, c=c, gamma=gamma, tol=tol, q=q, eigenvalues=vals, eigenvectors=vecs)
        return vals, vecs
