"""Generate tests/golden/f3.npz from the UNMODIFIED reference: ssl.centered_kernel (graphlearning/ssl.py:1345-1424) and
clustering.incres (graphlearning/clustering.py:283-371) on the two-moons and the 2000-node blob graphs.  Both draw from
numpy's global random stream; the seed set before each call is stored with the result.
    python -m oracle.make_golden_f3        TEST INFRASTRUCTURE ONLY."""
import importlib
import os

import numpy as np
from scipy import sparse

from ._refimport import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    gl = load_reference()
    clustering = importlib.import_module("gl_ref.clustering")
    out = {}
    for name, fn, tkey in (("moons", "twomoons500.npz", "train_ind"), ("blobs", "blobs2000.npz", "train_ind5")):
        g = np.load(os.path.join(OUT, fn))
        W = sparse.csr_matrix((g["W_data"], g["W_indices"], g["W_indptr"]), shape=tuple(g["W_shape"]))
        labels, ti = g["labels"], g[tkey]
        np.random.seed(11)
        m = gl.ssl.centered_kernel(W)
        out[name + "_ck_u"] = np.array(m.fit(ti, labels[ti]))
        out[name + "_ck_pred"] = np.array(m.predict())
        print(name, "centered kernel accuracy %.2f" % gl.ssl.ssl_accuracy(out[name + "_ck_pred"], labels, ti))
        if name == "moons":                  # INCRES needs a connected graph (its grow loop runs until F > 0 everywhere)
            np.random.seed(5)
            out["moons_incres"] = np.array(clustering.incres(W, 2, T=30).fit_predict())
    # a connected 3-cluster graph for INCRES: three overlapping Gaussian clouds in the plane, k = 12
    rng = np.random.default_rng(2)
    X = np.concatenate([rng.normal(size=(500, 2)) + c for c in ((0, 0), (3.2, 0), (1.6, 2.8))])
    W3 = gl.weightmatrix.knn(X, 12)
    from scipy.sparse.csgraph import connected_components
    assert connected_components(W3)[0] == 1
    out["clouds_W_data"], out["clouds_W_indices"], out["clouds_W_indptr"] = W3.data, W3.indices, W3.indptr
    out["clouds_labels"] = np.repeat(np.arange(3), 500)
    np.random.seed(7)
    out["clouds_incres"] = np.array(clustering.incres(W3, 3, T=40).fit_predict())
    print("incres accuracy: moons %.2f clouds %.2f" % (clustering.clustering_accuracy(out["moons_incres"], np.load(os.path.join(OUT, "twomoons500.npz"))["labels"]),
                                                        clustering.clustering_accuracy(out["clouds_incres"], out["clouds_labels"])))
    np.savez_compressed(os.path.join(OUT, "f3.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
