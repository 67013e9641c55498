"""Graph-based semi-supervised learning on the B200 backend: mirror of the hot-path classes of the reference
graphlearning/ssl.py (base class :131-511, poisson :513-693, laplace :1106-1261).

Same constructor kwargs, same fit/predict/fit_predict contract ((n,c) float64 scores in, int labels out).
Reference errors that were sys.exit(...) strings are ValueError / RuntimeError here.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
from scipy import sparse

from . import _lib, graph, utils


results_dir = os.path.join(os.getcwd(), "results")          # ssl.py:129


class ssl:
    """Base class.  Reference graphlearning/ssl.py:131-290, 439-481."""

    def __init__(self, W, class_priors):
        if W is None:
            self.graph = None
        else:
            self.set_graph(W)
        self.prob = None
        self.fitted = False
        self.name = ""
        self.accuracy_filename = ""
        self.requires_eig = False
        self.onevsrest = False
        self.similarity = True
        self.class_priors = class_priors
        if self.class_priors is not None:
            self.class_priors = self.class_priors / np.sum(self.class_priors)
        self.weights = 1
        self.class_priors_error = 1

    def set_graph(self, W):
        self.graph = W if type(W) == graph.graph else graph.graph(W)

    def volume_label_projection(self):
        """ssl.py:172-209: up to 10^4 rounds of predict -> class sizes -> weight update, in ONE kernel launch on the device
        (csrc/mbo.cu, glb_volume_projection): the same fp64 operations in the same order, so `weights`, the labels and
        `class_priors_error` are the reference's."""
        k = self.prob.shape[1]
        if type(self.weights) == int:
            self.weights = np.ones((k,))
        labels, self.weights, self.class_priors_error, self.projection_rounds = project_labels(
            self.prob, self.class_priors, self.weights, self.similarity)
        return labels

    def predict(self, ignore_class_priors=False):
        """ssl.py:230-266: global min-max scaling, then argmax (first maximum wins)."""
        if not self.fitted:
            raise RuntimeError("Model has not been fitted yet.")
        w = 1 if ignore_class_priors else self.weights
        scores = self.prob - np.min(self.prob)
        scores = scores / np.max(scores)
        if self.similarity:
            return np.argmax(scores * w, axis=1)
        return np.argmin(scores * w, axis=1)

    def fit_predict(self, train_ind, train_labels, all_labels=None):
        self.fit(train_ind, train_labels, all_labels=all_labels)
        return self.predict()

    def get_accuracy_filename(self):
        """ssl.py:212-227: `<accuracy_filename>[_classpriors]_accuracy.csv`."""
        return self.accuracy_filename + ("_classpriors" if self.class_priors is not None else "") + "_accuracy.csv"

    def ssl_trials(self, trainsets, labels, num_cores=1, tag="", save_results=True, overwrite=False, num_trials=-1):
        """Fit on every training set of a list and record `number of labels, accuracy[, with priors, priors error]` per trial:
        on the screen and, with save_results, as results/<tag><get_accuracy_filename()> - the file format of the reference's
        harness (ssl.py:292-396), which its trials_statistics / accuracy tables read.  The reference spreads the trials
        over `num_cores` joblib workers, each rebuilding the graph state; here the trials run one after the other on the GPU
        against the device state cached on the graph object (poisson_handle etc.), so `num_cores` is accepted and unused."""
        import os
        trainsets = list(trainsets)[:num_trials] if num_trials > 0 else list(trainsets)
        with_priors = self.class_priors is not None
        header = "Number of labels,Accuracy" + (",Accuracy with class priors,Class priors error" if with_priors else "")
        print("\nModel: " + self.name)
        outfile = None
        if save_results:
            os.makedirs(results_dir, exist_ok=True)
            outfile = os.path.join(results_dir, tag + self.get_accuracy_filename())
            if os.path.exists(outfile) and not overwrite:
                print("Aborting: SSL trial (" + self.get_accuracy_filename() + ") already completed , and overwrite is False.")
                return
            with open(outfile, "w") as f:
                f.write(header + "\n")
            print("Results File: " + outfile)
        print("\n" + header)
        for train_ind in trainsets:
            train_ind = np.asarray(train_ind)
            pred = self.fit_predict(train_ind, labels[train_ind])
            acc = ssl_accuracy(pred, labels, train_ind)
            if with_priors:
                plain = ssl_accuracy(self.predict(ignore_class_priors=True), labels, train_ind)
                line = "%d,%.2f,%.2f,%.5f" % (len(train_ind), plain, acc, self.class_priors_error)
            else:
                line = "%d,%.2f" % (len(train_ind), acc)
            print(line)
            if outfile:
                with open(outfile, "a+") as f:
                    f.write(line + "\n")

    def trials_statistics(self, tag=""):
        """(label counts, mean accuracy, std of accuracy, trials per label count) from the csv of ssl_trials (ssl.py:398-437)."""
        import os
        X = np.atleast_2d(np.loadtxt(os.path.join(results_dir, tag + self.get_accuracy_filename()), delimiter=",", skiprows=1))
        counts = np.unique(X[:, 0])
        mean = np.array([np.mean(X[X[:, 0] == m, 1:], axis=0) for m in counts])
        std = np.array([np.std(X[X[:, 0] == m, 1:], axis=0) for m in counts])
        return counts, mean, std, int(len(X[:, 0]) / len(counts))

    def fit(self, train_ind, train_labels, all_labels=None):
        """ssl.py:439-481."""
        if self.graph is None:
            raise RuntimeError("SSL object has no graph. Use set_graph() to provide a graph for SSL.")
        self.fitted = True
        train_ind = np.asarray(train_ind)
        train_labels = np.asarray(train_labels)
        if self.onevsrest:
            unique_labels = np.unique(train_labels)
            if hasattr(self, "_fit_onevsrest") and 1 < len(unique_labels) <= 32:
                # the classes share the graph and the labelled rows: one batched solve, column i identical to
                # self._fit(train_ind, train_labels == unique_labels[i]) of the reference loop (ssl.py:469-474)
                self.prob = self._fit_onevsrest(train_ind, train_labels[:, None] == unique_labels[None, :])
            else:
                self.prob = np.zeros((self.graph.num_nodes, len(unique_labels)))
                for i, l in enumerate(unique_labels):
                    self.prob[:, i] = self._fit(train_ind, train_labels == l)
        else:
            self.prob = self._fit(train_ind, train_labels, all_labels=all_labels)
        if self.class_priors is not None:
            self.volume_label_projection()
        return self.prob

    def _fit(self, train_ind, train_labels, all_labels=None):
        raise NotImplementedError("Must override _fit")


def _as_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def project_labels(prob, class_priors, weights, similarity=True, device_prob=None):
    """Volume-constrained label projection on the device (glb_volume_projection; reference ssl.py:172-209).
    prob: (n,k) float64 host array, or device_prob = (torch tensor (n, ld) float64, k) already in HBM.
    Returns (labels int64 (n,), weights (k,), max |class size - prior|, rounds)."""
    from . import device
    torch = device._torch()
    if device_prob is None:
        P = torch.from_numpy(np.ascontiguousarray(prob, dtype=np.float64)).cuda()
        n, k = P.shape
        ld = k
    else:
        P, k = device_prob
        n, ld = P.shape
    pri = torch.from_numpy(np.ascontiguousarray(class_priors, dtype=np.float64)).cuda()
    w = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64).copy()).cuda()
    labels = torch.empty(n, dtype=torch.int64, device="cuda")
    err = torch.zeros(1, dtype=torch.float64, device="cuda")
    rounds = torch.zeros(1, dtype=torch.int32, device="cuda")
    _lib.call("glb_volume_projection", device.ptr(P), n, int(k), int(ld), device.ptr(pri), int(bool(similarity)), 10000, 1e-3,
              device.ptr(w), device.ptr(labels), device.ptr(err), device.ptr(rounds), device.cur_stream())
    return labels.cpu().numpy(), w.cpu().numpy(), float(err.item()), int(rounds.item())


class poisson(ssl):
    """Poisson learning.  Reference graphlearning/ssl.py:513-693.

    solver='gradient_descent' runs the fused CSR-SpMM iterate on the GPU (poisson.cu) through the
    host-buffer entry point glb_poisson_gd_host; `use_cuda` is accepted for API compatibility (the GPU
    is always used).  Attribute `iterations` holds the iteration count T of the last fit.
    """

    def __init__(self, W=None, class_priors=None, solver="conjugate_gradient", p=1, use_cuda=False, min_iter=50,
                 max_iter=1000, tol=1e-3, spectral_cutoff=10):
        super().__init__(W, class_priors)
        if solver not in ["conjugate_gradient", "spectral", "gradient_descent"]:
            raise ValueError("Invalid Poisson solver")
        self.solver = solver
        self.p = p
        if p != 1:
            self.solver = "spectral"
        self.use_cuda = use_cuda
        self.min_iter = min_iter
        self.max_iter = max_iter
        self.tol = tol
        self.spectral_cutoff = spectral_cutoff
        self.iterations = None
        self.gpu_launches = 0
        fname = "_poisson"
        if self.p != 1:
            fname += "_p%.2f" % p
        if self.solver == "spectral":
            fname += "_N%d" % self.spectral_cutoff
        self.accuracy_filename = fname
        self.name = "Poisson Learning"

    def _source(self, train_ind, train_labels):
        """ssl.py:611-622."""
        n = self.graph.num_nodes
        k = len(np.unique(train_labels))
        onehot = utils.labels_to_onehot(train_labels, k)
        source = np.zeros((n, onehot.shape[1]))
        source[train_ind] = onehot - np.mean(onehot, axis=0)
        return source, k

    def _fit_gd_verbose(self, source, train_ind, all_labels):
        """all_labels given: the reference prints the accuracy after EVERY iteration (ssl.py:672-677).  Same iterate, one
        launch per iteration (glb_poisson_iterate with T = 1) and a download of the scores for predict() in between."""
        from . import device
        op = device.PoissonOperator(self.graph.weight_matrix)
        T = op.mixing_T(train_ind, self.min_iter, self.max_iter)
        Db = op.source_to_Db(source)
        c = source.shape[1]
        u = device._torch().zeros_like(Db)
        nl = 0
        for t in range(1, T + 1):
            u, l = op.iterate(Db, 1, u0=u, c=c)
            nl += l
            self.prob = op.unpack(u, c).cpu().numpy()
            print("%d,Accuracy = %.2f" % (t, ssl_accuracy(self.predict(), all_labels, train_ind)))
        self.iterations = T
        self.gpu_launches = nl
        return op.unpack(u, c).cpu().numpy()

    def _fit(self, train_ind, train_labels, all_labels=None):
        W = self.graph.weight_matrix
        n = self.graph.num_nodes
        if self.solver == "gradient_descent" and all_labels is None:
            # the source term is zero outside the labelled rows (ssl.py:619-622): only those rows go to the device
            k = len(np.unique(train_labels))
            onehot = utils.labels_to_onehot(train_labels, k)
            if onehot.shape[1] != k:
                # the reference adds an (n,k) array to an (n,width) one here and fails in numpy broadcasting
                raise ValueError("train_labels must be 0..k-1 for the gradient_descent solver")
            u, T, nl = self.graph.poisson_handle().fit_rows(train_ind, onehot - np.mean(onehot, axis=0), train_ind,
                                                            self.min_iter, self.max_iter)
            self.iterations = T
            self.gpu_launches = nl
            return u
        source, k = self._source(train_ind, train_labels)
        if self.solver == "gradient_descent":
            if source.shape[1] != k:
                raise ValueError("train_labels must be 0..k-1 for the gradient_descent solver")
            return self._fit_gd_verbose(source, train_ind, all_labels)
        if self.solver == "conjugate_gradient":
            # ssl.py:624-629: u = D^-1/2 conjgrad(L_normalized, D^-1/2 source, tol)
            if source.shape[1] != k:
                raise ValueError("train_labels must be 0..k-1")
            W0 = sparse.csr_matrix(W - sparse.spdiags(W.diagonal(), 0, n, n))          # :615-616
            G = graph.graph(W0)
            L = G.laplacian(normalization="normalized")
            D = G.degree_matrix(p=-0.5)
            v, (it, err, nl) = utils.conjgrad(L, D * source, tol=self.tol, return_info=True)
            self.iterations = it
            self.gpu_launches = nl
            return D * v
        # ssl.py:680-688: spectral solver on the leading eigenvectors of the random-walk Laplacian
        W0 = sparse.csr_matrix(W - sparse.spdiags(W.diagonal(), 0, n, n))              # :615-617
        G = graph.graph(W0)
        vals, vecs = G.eigen_decomp(normalization="randomwalk", k=self.spectral_cutoff + 1)
        self.gpu_launches = G.gpu_launches
        V = vecs[:, 1:]
        vals = vals[1:]
        if self.p != 1:
            vals = vals ** self.p
        L = sparse.spdiags(1 / vals, 0, self.spectral_cutoff, self.spectral_cutoff)
        return V @ (L @ (V.T @ source))


class poisson_mbo(ssl):
    """PoissonMBO.  Reference graphlearning/ssl.py:696-839: Poisson learning as the initial guess, then T rounds of Ns heat
    steps u <- (I - dt L) u + mu dt source (the same CSR x dense-label-matrix product as the Poisson iterate, here in fp64
    like the reference's CPU branch :822-823, on the block SpMM of spectral.cu) followed by the volume-constrained
    projection (device, csrc/mbo.cu).  The label matrix stays in HBM for the whole loop; only the result comes back.
    `use_cuda` is accepted for API compatibility: the GPU is always used."""

    def __init__(self, W=None, class_priors=None, solver="conjugate_gradient", use_cuda=False, min_iter=50, max_iter=1000,
                 tol=1e-3, spectral_cutoff=10, Ns=40, mu=1, T=20):
        super().__init__(W, class_priors)
        if class_priors is None:
            raise ValueError("PoissonMBO needs class_priors")
        self.poisson_model = poisson(W, solver=solver, use_cuda=use_cuda, min_iter=min_iter, max_iter=max_iter, tol=tol,
                                     spectral_cutoff=spectral_cutoff)
        self.Ns, self.mu, self.T, self.use_cuda = Ns, mu, T, use_cuda
        fname = "_poisson_mbo"
        if solver == "spectral":
            fname += "_N%d" % spectral_cutoff
        fname += "_Ns_%d_mu_%.2f_T_%d" % (Ns, mu, T)
        self.accuracy_filename = fname
        self.name = "Poisson MBO"
        self.gpu_launches = 0

    def _fit(self, train_ind, train_labels, all_labels=None):
        from . import device, spectral
        torch = device._torch()
        Ns, mu, T = self.Ns, self.mu, self.T
        n = self.graph.num_nodes
        k = len(np.unique(train_labels))
        W = self.graph.weight_matrix
        W = sparse.csr_matrix(W - sparse.spdiags(W.diagonal(), 0, n, n))               # :793-795
        G = graph.graph(W)
        onehot = utils.labels_to_onehot(train_labels, k)
        source = np.zeros((n, onehot.shape[1]))
        source[train_ind] = onehot - np.mean(onehot, axis=0)                            # :798-800
        labels = self.poisson_model.fit_predict(train_ind, train_labels, all_labels=all_labels)      # :803
        dt = 1 / np.max(G.degree_vector())                                              # :807
        P = sparse.csr_matrix(sparse.identity(n) - dt * G.laplacian())                  # :810
        ops = spectral.BlockOps(P)
        Db = ops.upload(mu * dt * source)                                               # :811
        u = ops.upload(utils.labels_to_onehot(labels, k))                               # :804
        tmp = ops.new(k)
        ld = int(u.shape[1])
        pri = torch.from_numpy(np.ascontiguousarray(self.class_priors, dtype=np.float64)).cuda()
        if type(self.weights) == int:
            self.weights = np.ones((k,))
        w = torch.from_numpy(np.ascontiguousarray(self.weights, dtype=np.float64).copy()).cuda()
        lab = torch.empty(n, dtype=torch.int64, device="cuda")
        err = torch.zeros(1, dtype=torch.float64, device="cuda")
        rounds = torch.zeros(1, dtype=torch.int32, device="cuda")
        for i in range(T):
            for _ in range(Ns):                                                         # heat steps, :822-823
                ops.spmm(u, k, out=tmp, Y2=Db, gamma=1.0)
                u, tmp = tmp, u
            # projection step, :826-829 (weights persist from round to round, as self.weights does in the reference)
            _lib.call("glb_volume_projection", device.ptr(u), n, k, ld, device.ptr(pri), int(bool(self.similarity)), 10000, 1e-3,
                      device.ptr(w), device.ptr(lab), device.ptr(err), device.ptr(rounds), device.cur_stream())
            _lib.call("glb_onehot_f64", device.ptr(lab), n, k, ld, device.ptr(u), device.cur_stream())
            if all_labels is not None:
                print("%d, Accuracy = %.2f" % (i, ssl_accuracy(lab.cpu().numpy(), all_labels, train_ind)))
        self.weights = w.cpu().numpy()
        self.class_priors_error = float(err.item())
        self.gpu_launches = ops.launches + 2 * T
        return u[:, :k].cpu().numpy()


class laplace(ssl):
    """Laplace learning (label propagation).  Reference graphlearning/ssl.py:1106-1261.

    Solves tau u + L u = 0 on the unlabelled nodes with the labels as Dirichlet data: the Jacobi-scaled
    sub-system M A M v = M b (ssl.py:1222-1246) is assembled on the device (laplace.cu) and solved by the
    multi-column CG (cg.cu) without leaving HBM; `system` restates the reference's scipy assembly (order > 1).  reweighting in {'poisson', 'wnll', 'properly'} goes through
    graph.reweight as in the reference (ssl.py:1209-1214).
    """

    def __init__(self, W=None, class_priors=None, X=None, reweighting="none", normalization="combinatorial", tau=0,
                 order=1, mean_shift=False, tol=1e-5, alpha=2, zeta=1e7, r=0.1):
        super().__init__(W, class_priors)
        self.reweighting = reweighting
        self.normalization = normalization
        self.mean_shift = mean_shift
        self.tol = tol
        self.order = order
        self.X = X
        if type(tau) in [float, int]:
            self.tau = np.ones(self.graph.num_nodes) * tau
        elif type(tau) is np.ndarray:
            self.tau = tau
        else:
            raise ValueError("tau must be a number or a numpy array")
        self.iterations = None
        self.gpu_launches = 0
        fname = "_laplace"
        self.name = "Laplace Learning"
        if self.reweighting != "none":
            fname += "_" + self.reweighting
            self.name += ": " + self.reweighting + " reweighted"
        if self.normalization != "combinatorial":
            fname += "_" + self.normalization
            self.name += " " + self.normalization
        if self.mean_shift:
            fname += "_meanshift"
            self.name += " with meanshift"
        if self.order > 1:
            fname += "_order%d" % int(self.order)
            self.name += " order %d" % int(self.order)
        if np.max(self.tau) > 0:
            fname += "_tau_%.3f" % np.max(self.tau)
            self.name += " tau=%.3f" % np.max(self.tau)
        self.accuracy_filename = fname

    def system(self, train_ind, train_labels):
        """(M A M, M b, M, idx, F): the linear system of ssl.py:1208-1246 (reweighting :1209-1214)."""
        if self.reweighting == "none":
            G = self.graph
        else:
            G = graph.graph(self.graph.reweight(train_ind, method=self.reweighting, normalization=self.normalization, X=self.X))
        n = G.num_nodes
        k = len(np.unique(train_labels))
        L = sparse.spdiags(self.tau, 0, n, n) + G.laplacian(normalization=self.normalization)
        Lp = L
        for _ in range(1, int(self.order)):
            Lp = L * Lp
        L = sparse.csr_matrix(Lp)
        F = utils.labels_to_onehot(train_labels, k)
        idx = np.full((n,), True, dtype=bool)
        idx[train_ind] = False
        b = (-L[:, train_ind] * F)[idx, :]
        A = L[idx, :][:, idx]
        m = A.shape[0]
        M = sparse.spdiags(1 / np.sqrt(A.diagonal() + 1e-10), 0, m, m).tocsr()
        return sparse.csr_matrix(M * A * M), M * b, M, idx, F

    def _fit(self, train_ind, train_labels, all_labels=None):
        n = self.graph.num_nodes
        train_ind = np.asarray(train_ind)
        fused = (int(self.order) == 1 and self.normalization in ("combinatorial", "randomwalk", "normalized")
                 and len(np.unique(train_ind)) == len(train_ind) < n)
        if fused:
            u = self._fit_device(train_ind, train_labels)
        else:
            # order > 1 (powers of L are sparse-sparse products), repeated labelled nodes: the reference's scipy assembly,
            # then the CG on the device
            MAM, Mb, M, idx, F = self.system(train_ind, train_labels)
            v, (it, err, nl) = utils.conjgrad(MAM, Mb, tol=self.tol, return_info=True)
            self.iterations = it
            self.gpu_launches = nl
            u = np.zeros((n, F.shape[1]))
            u[idx, :] = M * v
            u[train_ind, :] = F
        if self.mean_shift:
            u -= np.mean(u, axis=0)
        return u

    def _fit_device(self, train_ind, train_labels):
        """ssl.py:1208-1255 in one call of glb_laplace_graph_fit: Laplacian, tau, Dirichlet sub-system, Jacobi scaling, CG and
        the scatter of the solution all stay in HBM; the host computes the n values of d^p (numpy, as the reference)."""
        import ctypes
        from . import _lib
        if self.reweighting == "none":
            G = self.graph
        else:
            G = graph.graph(self.graph.reweight(train_ind, method=self.reweighting, normalization=self.normalization, X=self.X))
        n = G.num_nodes
        k = len(np.unique(train_labels))
        F = np.ascontiguousarray(utils.labels_to_onehot(train_labels, k), dtype=np.float64)
        tau = np.ascontiguousarray(np.broadcast_to(self.tau, (n,)), dtype=np.float64) if np.any(self.tau != 0) else None
        ti = np.ascontiguousarray(np.where(train_ind < 0, train_ind + n, train_ind), dtype=np.int64)
        # W and the scalings stay in HBM across the fits on this graph (a reweighted graph is new for every labelled set)
        u, it, err, nl, ms = G.laplace_handle(self.normalization, tau).fit(ti, F, self.tol)
        self.iterations = it
        self.gpu_launches = nl
        self.cg_info = {"device_ms": float(ms[0]), "system_nnz": int(ms[1]), "unknowns": int(ms[2])}
        return u


class amle(ssl):
    """AMLE learning, one-vs-rest.  Reference graphlearning/ssl.py:1569-1614; sweeps on the GPU (plaplace.cu)."""

    def __init__(self, W=None, class_priors=None, tol=1e-3, max_num_it=1e5, weighted=False, prog=False):
        super().__init__(W, class_priors)
        self.tol = tol
        self.max_num_it = max_num_it
        self.weighted = weighted
        self.prog = prog
        self.onevsrest = True
        self.accuracy_filename = "_amle"
        if not self.weighted:
            self.accuracy_filename += "_unweighted"
        self.name = "AMLE"

    def _fit(self, train_ind, train_labels, all_labels=None):
        return self.graph.amle(train_ind, train_labels, tol=self.tol, max_num_it=self.max_num_it,
                               weighted=self.weighted, prog=self.prog)

    def _fit_onevsrest(self, train_ind, onehot):
        return self.graph.amle(train_ind, np.asarray(onehot, dtype=np.float64), tol=self.tol, max_num_it=self.max_num_it,
                               weighted=self.weighted, prog=self.prog)


class plaplace(ssl):
    """Graph p-Laplace classifier, one-vs-rest.  Reference graphlearning/ssl.py:1681-1727."""

    def __init__(self, W=None, class_priors=None, p=10, max_num_it=1e6, tol=1e-1, fast=True):
        super().__init__(W, class_priors)
        self.p = p
        self.max_num_it = max_num_it
        self.tol = tol
        self.onevsrest = True
        self.fast = fast
        if fast:
            self.tol = 1e-5
        self.accuracy_filename = "_plaplace_p%.2f" % self.p
        self.name = "p-Laplace (p=%.2f)" % self.p

    def _fit(self, train_ind, train_labels, all_labels=None):
        return self.graph.plaplace(train_ind, train_labels, self.p, max_num_it=self.max_num_it, tol=self.tol,
                                   fast=self.fast)

    def _fit_onevsrest(self, train_ind, onehot):
        if not self.fast:                                         # the Jacobi barrier-function solver stays one class at a time
            return np.stack([self._fit(train_ind, onehot[:, k]) for k in range(onehot.shape[1])], axis=1)
        return self.graph.plaplace(train_ind, np.asarray(onehot, dtype=np.float64), self.p, max_num_it=self.max_num_it,
                                   tol=self.tol, fast=True)


class randomwalk(ssl):
    """Lazy random walk classification.  Reference graphlearning/ssl.py:1731-1793: one Jacobi-scaled multi-column CG
    solve of ((1 - alpha) I + alpha L_normalized) u = Y, on the GPU (cg.cu) through utils.conjgrad."""

    def __init__(self, W=None, class_priors=None, alpha=0.95):
        super().__init__(W, class_priors)
        self.alpha = alpha
        self.accuracy_filename = "_randomwalk"
        self.name = "Lazy Random Walks"
        self.iterations = None
        self.gpu_launches = 0

    def _fit(self, train_ind, train_labels, all_labels=None):
        alpha = self.alpha
        n = self.graph.num_nodes
        W = self.graph.weight_matrix
        W = W - sparse.spdiags(W.diagonal(), 0, n, n)
        G = graph.graph(W)
        L = (1 - alpha) * sparse.identity(n) + alpha * G.laplacian(normalization="normalized")
        m = L.shape[0]
        M = sparse.spdiags(1 / np.sqrt(L.diagonal() + 1e-10), 0, m, m).tocsr()
        k = len(np.unique(train_labels))
        onehot = utils.labels_to_onehot(train_labels, k)
        Y = np.zeros((n, onehot.shape[1]))
        Y[train_ind, :] = onehot
        u, (it, err, nl) = utils.conjgrad(M * L * M, M * Y, tol=1e-6, return_info=True)
        self.iterations, self.gpu_launches = it, nl
        return M * u


class centered_kernel(ssl):
    """Centered kernel method of Mai & Couillet.  Reference graphlearning/ssl.py:1345-1424: a power iteration for the largest
    eigenvalue of the doubly centred weight matrix A = H W H (H = I - 11^T / n), then the fixed point
    u <- (1/alpha) A u with the labelled rows held fixed, until max |change| <= tol.  A is never formed: with d = W 1 and
    s = 1^T x, H W H x = y - 1 mean(y)^T, y = W x - d s^T / n, i.e. one fp64 block SpMM with the rank-one term fused into
    its epilogue (spectral.cu), one column-sum reduction and one fused update kernel (mbo.cu) per iteration; the label
    matrix stays in HBM.  The start vector comes from numpy's global stream exactly as in the reference (:1399)."""

    def __init__(self, W=None, class_priors=None, tol=1e-10, power_it=100, alpha=1.05):
        super().__init__(W, class_priors)
        self.tol = tol
        self.power_it = power_it
        self.alpha = alpha
        self.accuracy_filename = "_centered_kernel"
        self.name = "Centered Kernel"
        self.iterations = None
        self.gpu_launches = 0

    def _fit(self, train_ind, train_labels, all_labels=None):
        from . import device, spectral
        torch = device._torch()
        n = self.graph.num_nodes
        k = len(np.unique(train_labels))
        W = self.graph.weight_matrix
        W = sparse.csr_matrix(W - sparse.spdiags(W.diagonal(), 0, n, n))               # :1385-1386
        onehot = utils.labels_to_onehot(train_labels, k)
        kk = onehot.shape[1]
        K = np.zeros((n, kk))
        K[train_ind] = onehot
        K[train_ind, :] -= np.sum(K, axis=0) / len(train_ind)                           # :1389-1393
        ops = spectral.BlockOps(W)
        d = W * np.ones(n)
        ones = ops.upload(np.ones((n, 1)))
        nl = 0

        def centred_product(x, c, dcols, y):
            """y = W (x - 1 mean(x)^T) = W x - d s^T / n  and the column means of y"""
            nonlocal nl
            s = ops.gram(ones, 1, x, c)[0]
            ops.spmm(x, c, out=y, Y1=dcols, beta=1.0, bcol=-(1 / n) * s)
            t = ops.gram(ones, 1, y, c)[0]
            nl += 5
            return (1 / n) * t

        # largest eigenvalue of A by power iteration (:1399-1404)
        e = ops.upload(np.random.rand(n, 1))
        e2 = ops.new(1)
        d1 = ops.upload(d[:, None])
        y = ops.new(1)
        lam = 0.0
        for _ in range(self.power_it):
            m = centred_product(e, 1, d1, y)
            ee = ops.gram(e, 1, e, 1)[0, 0]
            # w = y - mean(y);  e^T w = e^T y - mean(y) sum(e);  |w|^2 = y^T y - n mean(y)^2
            ey = ops.gram(e, 1, y, 1)[0, 0]
            se = ops.gram(ones, 1, e, 1)[0, 0]
            yy = ops.gram(y, 1, y, 1)[0, 0]
            lam = abs((ey - m[0] * se) / ee)
            wn = np.sqrt(max(yy - n * m[0] * m[0], 0.0))
            # e = (y - mean(y)) / |w|   (the epilogue of the block kernel as a linear combination: alpha = 0)
            ops.spmm(e, 1, out=e2, alpha=0.0, Y1=y, beta=1.0 / wn, Y2=ones, gamma=-m[0] / wn)
            e, e2 = e2, e
            nl += 9
        # fixed point (:1407-1413)
        alpha = self.alpha * lam
        u = ops.upload(K)
        yk = ops.new(kk)
        dk = ops.upload(np.repeat(d[:, None], kk, axis=1))
        mask = np.zeros(n, dtype=np.uint8)
        mask[train_ind] = 1
        mask_d = torch.from_numpy(mask).cuda()
        mean_d = torch.empty(kk, dtype=torch.float64, device="cuda")
        err, it = 1.0, 0
        e_host = ctypes.c_double(0.0)
        while err > self.tol:
            m = centred_product(u, kk, dk, yk)
            mean_d.copy_(torch.from_numpy(np.ascontiguousarray(m)))
            _lib.call("glb_centered_step_f64", device.ptr(yk), int(yk.shape[1]), device.ptr(mean_d), float(1 / alpha), device.ptr(u),
                      int(u.shape[1]), device.ptr(mask_d), n, kk, ctypes.byref(e_host), device.cur_stream())
            err = e_host.value
            it += 1
            nl += 1
            if all_labels is not None:
                self.prob = u[:, :kk].cpu().numpy()
                print("%d,Accuracy = %.2f" % (it, ssl_accuracy(self.predict(), all_labels, train_ind)))
        self.iterations = it
        self.eigenvalue = lam
        self.gpu_launches = nl
        return u[:, :kk].cpu().numpy()


def ssl_accuracy(pred_labels, true_labels, train_ind):
    """Accuracy over the unlabelled points, in percent.  Reference graphlearning/ssl.py:1795-1834."""
    pred_labels = np.asarray(pred_labels)
    true_labels = np.asarray(true_labels)
    mask = np.ones(len(pred_labels), dtype=bool)
    if type(train_ind) != np.ndarray:
        print("Warning: ssl_accuracy has been updated and now requires the user to provide the indices of the "
              "labeled points, and not just the number of labels.")
    else:
        mask[train_ind] = False
    p, t = pred_labels[mask], true_labels[mask]
    I = t >= 0
    return 100 * np.mean(p[I] == t[I])
