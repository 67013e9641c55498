"""Generate tests/golden/plaplace2000.npz: graph.plaplace / graph.amle / ssl.plaplace / ssl.amle of the
UNMODIFIED reference Python (/root/reference/graphlearning/graph.py:1177-1332, ssl.py:1569-1614, 1681-1727)
on the 2000-node blob graph of tests/golden/blobs2000.npz.  Run by hand in the build container:

    python -m oracle.make_golden_plaplace

TEST INFRASTRUCTURE ONLY.  The reference's CPython module `cextensions` is not built in the read-only checkout,
so a stand-in module with the same two entry points (c_code/cextensions.cpp:19-107: same argument order, arrays
updated in place, T and the flags passed as doubles) forwards to oracle/_ref/liblp_ref.so - the reference's own
c_code/lp_iterate.cpp compiled from where it lies with -O2 (its setup.py uses -Ofast, which licenses
reassociation; bit-level parity is defined against the IEEE build).
"""
import ctypes
import os
import sys
import types

import numpy as np
from scipy import sparse

from . import c_oracle
from ._refimport import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _pad(a, dtype, fill):
    out = np.empty(len(a) + 1, dtype=dtype)
    out[:-1] = a
    out[-1] = fill
    return out


def install_cextensions_shim():
    ref = c_oracle.ref()
    dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)
    d = lambda a: a.ctypes.data_as(dp)
    i = lambda a: a.ctypes.data_as(ip)
    mod = types.ModuleType("gl_ref.cextensions")

    def lp_iterate(uu, ul, II, J, W, ind, val, p, Td, tol, progd):
        n, M, m = len(uu), len(II), len(ind)
        II_, J_, W_ = _pad(II, np.int32, 0), _pad(J, np.int32, -1), _pad(W, np.float64, 0)
        ref.ref_lp_iterate(d(uu), d(ul), i(II_), i(J_), d(W_), i(ind), d(val), ctypes.c_double(p), ctypes.c_int(int(Td)),
                           ctypes.c_double(tol), ctypes.c_int(n), ctypes.c_int(M), ctypes.c_int(m))

    def lip_iterate(u, II, J, W, ind, val, Td, tol, progd, weightedd, alpha, beta):
        n, M, m = len(u), len(II), len(ind)
        II_, J_, W_ = _pad(II, np.int32, 0), _pad(J, np.int32, -1), _pad(W, np.float64, 0)
        if bool(weightedd):
            ref.ref_lip_iterate_weighted(d(u), i(II_), i(J_), d(W_), i(ind), d(val), ctypes.c_int(int(Td)),
                                         ctypes.c_double(tol), ctypes.c_int(n), ctypes.c_int(M), ctypes.c_int(m))
        else:
            ref.ref_lip_iterate(d(u), i(II_), i(J_), d(W_), i(ind), d(val), ctypes.c_int(int(Td)), ctypes.c_double(tol),
                                ctypes.c_int(n), ctypes.c_int(M), ctypes.c_int(m), ctypes.c_double(alpha),
                                ctypes.c_double(beta))

    mod.lp_iterate, mod.lip_iterate = lp_iterate, lip_iterate
    sys.modules["gl_ref.cextensions"] = mod
    sys.modules["gl_ref"].cextensions = mod


def main():
    gl = load_reference()
    install_cextensions_shim()
    g = np.load(os.path.join(OUT, "blobs2000.npz"))
    W = sparse.csr_matrix((g["W_data"], g["W_indices"], g["W_indptr"]), shape=tuple(g["W_shape"]))
    labels, ti = g["labels"], g["train_ind5"]
    G = gl.graph.graph(W)
    out = dict(cI=G.I, cJ=G.J, cV=G.V, train_ind=ti)
    val = (labels[ti] == 0).astype(np.float64)
    out["val"] = val
    out["pl_fast_p3"] = G.plaplace(ti, val, 3)                                   # lip_iterate_main, tol 1e-6
    out["pl_fast_p10_T40"] = G.plaplace(ti, val, 10, max_num_it=40)
    out["pl_slow_p3"] = G.plaplace(ti, val, 3, tol=1e-1, fast=False)             # lp_iterate_main
    out["pl_slow_p3_T101"] = G.plaplace(ti, val, 3, tol=1e-9, max_num_it=101, fast=False)
    out["amle_w"] = G.amle(ti, val, tol=1e-5, max_num_it=1000, weighted=True)   # lip_iterate_weighted_main
    out["amle_w_T25"] = G.amle(ti, val, tol=1e-5, max_num_it=25, weighted=True)
    out["amle_u"] = G.amle(ti, val, tol=1e-5, max_num_it=1000, weighted=False)
    # a directed graph with empty rows (the `u[I[start[i]]]` read of an empty row, lp_iterate.cpp:163,223)
    Wd = sparse.csr_matrix(W, copy=True).tolil()
    for r in (5, 700, 1500):
        Wd[r, :] = 0
    Wd = sparse.csr_matrix(Wd); Wd.eliminate_zeros()
    Gd = gl.graph.graph(Wd)
    out["dI"], out["dJ"], out["dV"] = Gd.I, Gd.J, Gd.V
    out["amle_w_directed_T30"] = Gd.amle(ti, val, tol=1e-9, max_num_it=30, weighted=True)
    out["amle_u_directed_T30"] = Gd.amle(ti, val, tol=1e-9, max_num_it=30, weighted=False)   # NaN on the empty rows
    m = gl.ssl.plaplace(W, p=3)
    out["ssl_plaplace_p3"] = np.array(m.fit(ti, labels[ti])); out["ssl_plaplace_p3_pred"] = np.array(m.predict())
    m = gl.ssl.amle(W)
    out["ssl_amle"] = np.array(m.fit(ti, labels[ti])); out["ssl_amle_pred"] = np.array(m.predict())
    np.savez_compressed(os.path.join(OUT, "plaplace2000.npz"), **out)
    print("plaplace2000 done:", {k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
