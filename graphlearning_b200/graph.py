"""Graph object (mirror of the hot-path part of reference graphlearning/graph.py:25-84,108-122,210-233,469-513).

Holds the weight matrix as a scipy CSR on the host - the same object the reference hands around - and
builds the degree / Laplacian matrices with the same scipy expressions; the arithmetic that matters for
the iterate (degrees, D^-1 scaling, transposition) is redone on the device from the CSR when a solver
runs (graph_ops.cu), so nothing here is on the timed path.
"""
from __future__ import annotations

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

    def laplacian(self, normalization="combinatorial", alpha=1):
        """graph.py:469-513."""
        I = sparse.identity(self.num_nodes)
        D = self.degree_matrix()
        if normalization == "combinatorial":
            L = D - self.weight_matrix
        elif normalization == "randomwalk":
            L = I - self.degree_matrix(p=-1) * self.weight_matrix
        elif normalization == "normalized":
            Dinv2 = self.degree_matrix(p=-0.5)
            L = I - Dinv2 * self.weight_matrix * Dinv2
        elif normalization == "coifmanlafon":
            D = self.degree_matrix(p=-alpha)
            L = graph(D * self.weight_matrix * D).laplacian(normalization="randomwalk")
        else:
            raise ValueError("Invalid option for graph Laplacian normalization.")
        return L.tocsr()
