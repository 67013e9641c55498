"""Clustering on the B200 backend: spectral clustering (the caller of graph.eigen_decomp for the 70k-node spectral
configuration) and INCRES (the reseeded power loop F <- P F).  Mirror of reference graphlearning/clustering.py:19-62 (base
class), 113-198 (spectral), 283-371 (incres) and the accuracy helpers (:470-550); k-means itself stays sklearn on the host
as in the reference (:196)."""
from __future__ import annotations

import numpy as np
import scipy.optimize as opt
import sklearn.cluster as cluster
from scipy import sparse

from . import graph


class clustering:
    def __init__(self, W, num_clusters):
        self.graph = W if type(W) == graph.graph else graph.graph(W)
        self.cluster_labels = None
        self.num_clusters = num_clusters
        self.fitted = False

    def predict(self):
        if not self.fitted:
            raise RuntimeError("Model has not been fitted yet.")
        return self.cluster_labels

    def fit_predict(self, all_labels=None):
        self.fit(all_labels=all_labels)
        return self.cluster_labels

    def fit(self, all_labels=None):
        self.fitted = True
        self.cluster_labels = self._fit(all_labels=all_labels)
        return self.cluster_labels

    def _fit(self, all_labels=None):
        raise NotImplementedError("Must override _fit")


class spectral(clustering):
    """clustering.py:113-198: 'combinatorial', 'ShiMalik' or 'NgJordanWeiss' spectral embedding + k-means."""

    def __init__(self, W, num_clusters, method="NgJordanWeiss", extra_dim=0):
        super().__init__(W, num_clusters)
        self.method = method
        self.extra_dim = extra_dim

    def embedding(self):
        n = self.graph.num_nodes
        k = self.num_clusters + self.extra_dim
        if self.method == "combinatorial":
            vals, vec = self.graph.eigen_decomp(k=k)
        elif self.method == "ShiMalik":
            vals, vec = self.graph.eigen_decomp(normalization="randomwalk", k=k)
        elif self.method == "NgJordanWeiss":
            vals, vec = self.graph.eigen_decomp(normalization="normalized", k=k)
            norms = np.sum(vec * vec, axis=1)
            T = sparse.spdiags(norms ** (-1 / 2), 0, n, n)
            vec = T @ vec
        else:
            raise ValueError("Invalid spectral clustering method " + str(self.method))
        return vec

    def _fit(self, all_labels=None):
        kmeans = cluster.KMeans(n_clusters=self.num_clusters).fit(self.embedding())
        return kmeans.labels_


class incres(clustering):
    """INCRES clustering (incremental reseeding).  Reference graphlearning/clustering.py:283-371: T rounds of plant (m random
    seeds per cluster, numpy's global stream as in the reference), grow (F <- P F with P = W D^-1 until every entry of F is
    positive) and harvest (argmax).  Grow and harvest run on the device: the fp64 block SpMM of spectral.cu, a min reduction
    and a row argmax (mbo.cu); only the n labels travel per round (the seeds are drawn on the host)."""

    def __init__(self, W, num_clusters, speed=5, T=200):
        super().__init__(W, num_clusters)
        self.speed = speed
        self.T = T
        self.gpu_launches = 0

    def _fit(self, all_labels=None):
        import ctypes
        from . import _lib, device, spectral
        torch = device._torch()
        n, speed, T, k = self.graph.num_nodes, self.speed, self.T, self.num_clusters
        Dm = np.maximum(int(speed * 1e-4 * n / k), 1)                                   # :337
        u = np.random.randint(0, k, size=n)                                             # :340
        J = np.arange(n).astype(int)
        D = self.graph.degree_matrix(p=-1)
        ops = spectral.BlockOps(sparse.csr_matrix(self.graph.weight_matrix * D))        # :347-348
        F = ops.new(k)
        G = ops.new(k)
        lab = torch.empty(n, dtype=torch.int64, device="cuda")
        lo = ctypes.c_double(0.0)
        nl = 0
        m = int(1)
        for i in range(T):
            Fh = np.zeros((n, k))                                                       # plant, :352-357
            for r in range(k):
                I = u == r
                ind = J[I]
                Fh[ind[np.random.choice(np.sum(I), m)], r] = 1
            F[:, :k] = torch.from_numpy(Fh).cuda()
            while True:                                                                 # grow, :360-361
                _lib.call("glb_min_nonneg_f64", device.ptr(F), n, k, int(F.shape[1]), ctypes.byref(lo), device.cur_stream())
                nl += 1
                if lo.value != 0:
                    break
                ops.spmm(F, k, out=G)
                F, G = G, F
            _lib.call("glb_argmax_rows_f64", device.ptr(F), n, k, int(F.shape[1]), device.ptr(lab), device.cur_stream())   # harvest, :364
            nl += 1
            u = lab.cpu().numpy()
            m = m + Dm
            if all_labels is not None:
                acc = clustering_accuracy(u, all_labels)
                print("Iteration " + str(i) + ": Accuracy = %.2f" % acc + "%%, #seeds= %d" % m)
        self.gpu_launches = nl + ops.launches
        return u


def clustering_accuracy(pred_labels, true_labels):
    """Accuracy in percent under the best matching of cluster ids to classes (Hungarian algorithm).
    Reference graphlearning/clustering.py:470-510."""
    pred_labels = np.asarray(pred_labels)
    true_labels = np.asarray(true_labels)
    unique_classes = np.unique(true_labels)
    unique_clusters = np.unique(pred_labels)
    C = np.zeros((len(unique_clusters), len(unique_classes)), dtype=float)
    for i, cl in enumerate(unique_clusters):
        for j, c in enumerate(unique_classes):
            C[i, j] = np.sum((pred_labels == cl) & (true_labels != c))
    row_ind, col_ind = opt.linear_sum_assignment(C)
    return 100 * (1 - C[row_ind, col_ind].sum() / len(pred_labels))


def purity(cluster_labels, true_labels):
    """Cluster purity: (overall purity in percent, fraction of the largest class per cluster).
    Reference graphlearning/clustering.py:513-550."""
    cluster_labels = np.asarray(cluster_labels)
    true_labels = np.asarray(true_labels)
    largest, size = [], []
    for cl in np.unique(cluster_labels):
        members = true_labels[cluster_labels == cl]
        largest.append(np.max(np.bincount(members)))
        size.append(len(members))
    largest, size = np.array(largest), np.array(size)
    return 100 * np.sum(largest) / np.sum(size), largest / size
