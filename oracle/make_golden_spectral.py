"""Generate tests/golden/spectral.npz: graph.eigen_decomp of the UNMODIFIED reference
(/root/reference/graphlearning/graph.py:623-806; ARPACK svds and utils.randomized_svd, utils.py:576-642) and the
spectral Poisson solver (ssl.py:680-688) on the two-moons and 2000-node blob graphs of the other fixtures.

    python -m oracle.make_golden_spectral

TEST INFRASTRUCTURE ONLY.  Eigenvectors are unique only up to sign (and rotation inside a repeated eigenvalue), so
the tests compare eigenvalues and spectral projectors, not vectors."""
import os

import numpy as np
from scipy import sparse

from ._refimport import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    gl = load_reference()
    out = {}
    m = np.load(os.path.join(OUT, "twomoons500.npz"))
    W = sparse.csr_matrix((m["W_data"], m["W_indices"], m["W_indptr"]), shape=tuple(m["W_shape"]))
    for norm in ("normalized", "randomwalk", "combinatorial"):
        vals, vecs = gl.graph.graph(W).eigen_decomp(normalization=norm, k=12)
        out["moons_%s_vals" % norm] = vals; out["moons_%s_vecs" % norm] = vecs
    np.random.seed(5)
    vals, vecs = gl.graph.graph(W).eigen_decomp(normalization="normalized", method="lowrank", k=8, c=30, q=20)
    out["moons_lowrank_vals"] = vals; out["moons_lowrank_vecs"] = vecs
    ti, labels = m["train_ind"], m["labels"]
    model = gl.ssl.poisson(W, solver="spectral", spectral_cutoff=10)
    out["moons_poisson_spectral"] = np.array(model.fit(ti, labels[ti])); out["moons_poisson_spectral_pred"] = np.array(model.predict())

    b = np.load(os.path.join(OUT, "blobs2000.npz"))
    Wb = sparse.csr_matrix((b["W_data"], b["W_indices"], b["W_indptr"]), shape=tuple(b["W_shape"]))
    # connect the ten blobs weakly so that the spectrum is simple enough for ARPACK to resolve (a 10-fold eigenvalue
    # 1 of the block-diagonal graph makes svds itself miss copies)
    rng = np.random.default_rng(3)
    r = rng.integers(0, 2000, 400); c = rng.integers(0, 2000, 400)
    E = sparse.csr_matrix((np.full(400, 0.05), (r, c)), shape=(2000, 2000))
    Wc = sparse.csr_matrix(Wb + E + E.T); Wc.setdiag(0); Wc.eliminate_zeros()
    out["blobs_Wc_data"], out["blobs_Wc_indices"], out["blobs_Wc_indptr"] = Wc.data, Wc.indices.astype(np.int32), Wc.indptr.astype(np.int32)
    vals, vecs = gl.graph.graph(Wc).eigen_decomp(normalization="normalized", k=30)
    out["blobs_normalized_vals"] = vals; out["blobs_normalized_vecs"] = vecs
    np.savez_compressed(os.path.join(OUT, "spectral.npz"), **out)
    print({k: v.shape for k, v in out.items()})
    print(out["blobs_normalized_vals"][:14])


if __name__ == "__main__":
    main()
