"""Generate tests/golden/randomwalk.npz: ssl.randomwalk of the UNMODIFIED reference (graphlearning/ssl.py:1731-1793) on the
two-moons and 2000-node blob graphs.    python -m oracle.make_golden_randomwalk        TEST INFRASTRUCTURE ONLY."""
import os

import numpy as np
from scipy import sparse

from ._refimport import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    gl = load_reference()
    out = {}
    for name, f, tkey in (("moons", "twomoons500.npz", "train_ind"), ("blobs", "blobs2000.npz", "train_ind5")):
        g = np.load(os.path.join(OUT, f))
        W = sparse.csr_matrix((g["W_data"], g["W_indices"], g["W_indptr"]), shape=tuple(g["W_shape"]))
        ti, labels = g[tkey], g["labels"]
        m = gl.ssl.randomwalk(W)
        out[name + "_u"] = np.array(m.fit(ti, labels[ti])); out[name + "_pred"] = np.array(m.predict())
    np.savez_compressed(os.path.join(OUT, "randomwalk.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
