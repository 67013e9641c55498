"""Generate tests/golden/mbo.npz from the UNMODIFIED reference: ssl.poisson_mbo (graphlearning/ssl.py:696-839),
ssl.volume_label_projection through ssl.laplace(class_priors=...) (:172-209, 476-477) and graph.page_rank
(graphlearning/graph.py:1374-1412) on the 2000-node blob graph and the two-moons graph.
    python -m oracle.make_golden_mbo        TEST INFRASTRUCTURE ONLY."""
import os

import numpy as np
from scipy import sparse

from ._refimport import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    gl = load_reference()
    out = {}
    g = np.load(os.path.join(OUT, "blobs2000.npz"))
    W = sparse.csr_matrix((g["W_data"], g["W_indices"], g["W_indptr"]), shape=tuple(g["W_shape"]))
    labels, ti = g["labels"], g["train_ind5"]
    priors = gl.utils.class_priors(labels)
    out["priors"] = priors
    # volume projection on given scores: Laplace learning with class priors
    m = gl.ssl.laplace(W, class_priors=priors)
    out["lap_prob"] = np.array(m.fit(ti, labels[ti]))
    out["lap_weights"] = np.array(m.weights); out["lap_pred"] = np.array(m.predict()); out["lap_err"] = np.array(m.class_priors_error)
    # projection of a deliberately unbalanced score matrix (many rounds)
    rng = np.random.default_rng(3)
    prob = rng.random((5000, 4)) * np.array([1.0, 1.3, 0.7, 1.1])
    m2 = gl.ssl.laplace(W, class_priors=np.array([0.4, 0.1, 0.3, 0.2]))
    m2.prob = prob.copy(); m2.fitted = True
    out["synth_prob"] = prob
    out["synth_pred"] = np.array(m2.volume_label_projection()); out["synth_weights"] = np.array(m2.weights)
    out["synth_err"] = np.array(m2.class_priors_error)
    # PoissonMBO (default conjugate-gradient initialisation; short run to keep the golden small)
    mb = gl.ssl.poisson_mbo(W, priors, Ns=20, T=6)
    out["mbo_u"] = np.array(mb.fit(ti, labels[ti])); out["mbo_pred"] = np.array(mb.predict()); out["mbo_weights"] = np.array(mb.weights)
    # page rank
    G = gl.graph.graph(W)
    out["pagerank"] = np.array(G.page_rank())
    v = np.zeros(W.shape[0]); v[ti] = 1 / len(ti)
    out["pagerank_personalised"] = np.array(G.page_rank(alpha=0.7, v=v, tol=1e-8))
    np.savez_compressed(os.path.join(OUT, "mbo.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
