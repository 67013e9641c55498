"""Spectral clustering on the B200 backend: the caller of graph.eigen_decomp for the 70k-node spectral configuration.
Mirror of reference graphlearning/clustering.py:19-62 (base class), 113-198 (spectral) and the accuracy helpers
(:470-550); k-means itself stays sklearn on the host as in the reference (:196)."""
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
    """Cluster purity in percent.  Reference graphlearning/clustering.py:513-550."""
    cluster_labels = np.asarray(cluster_labels)
    true_labels = np.asarray(true_labels)
    hits = 0
    for cl in np.unique(cluster_labels):
        members = true_labels[cluster_labels == cl]
        hits += np.max(np.bincount(members - members.min())) if len(members) else 0
    return 100 * hits / len(true_labels)
