"""Generate tests/golden/reweight.npz: graph.reweight and ssl.laplace(reweighting=...) of the UNMODIFIED reference
(/root/reference/graphlearning/graph.py:368-466, ssl.py:1209-1214) on the two-moons graph, plus the on-disk kNN format.

    python -m oracle.make_golden_reweight

TEST INFRASTRUCTURE ONLY."""
import os

import numpy as np
from scipy import sparse

from ._refimport import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    gl = load_reference()
    m = np.load(os.path.join(OUT, "twomoons500.npz"))
    W = sparse.csr_matrix((m["W_data"], m["W_indices"], m["W_indptr"]), shape=tuple(m["W_shape"]))
    ti, labels, X = m["train_ind"], m["labels"], m["X"]
    G = gl.graph.graph(W)
    out = {}
    for method, kw in (("poisson", {}), ("poisson", {"normalization": "normalized"}), ("wnll", {}), ("properly", {"X": X})):
        Wr = sparse.csr_matrix(G.reweight(ti, method=method, **kw))
        Wr.sort_indices()
        tag = method + ("_" + kw["normalization"] if "normalization" in kw else "")
        out["W_%s_data" % tag] = Wr.data; out["W_%s_indices" % tag] = Wr.indices; out["W_%s_indptr" % tag] = Wr.indptr
    for rw in ("poisson", "wnll"):
        model = gl.ssl.laplace(W, reweighting=rw)
        out["u_laplace_" + rw] = np.array(model.fit(ti, labels[ti])); out["p_laplace_" + rw] = np.array(model.predict())
    np.savez_compressed(os.path.join(OUT, "reweight.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
