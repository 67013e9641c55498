"""Time the p-Laplace / AMLE sweeps on the 70k-node benchmark graph: GPU (plaplace.cu through the C-ABI, host buffers
in and out) next to the plain-C oracle on one host core.  Prints one JSON object."""
import ctypes, json, os, sys, time
import numpy as np
from scipy import sparse
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import c_oracle, gl_oracle as orc
from graphlearning_b200 import _lib

n = 70000
X, labels = orc.synthetic_blobs(n, 8, c=10, seed=0)
ind, dist = orc.knnsearch(X.astype(np.float64), 11, method="kdtree")
W = sparse.csr_matrix(orc.knn_weights(ind, dist, 10))
I, J, V = orc.ccode_triplets(W)
ti = orc.one_per_class(labels, rate=5, seed=0).astype(np.int32)
val = (labels[ti] == 3).astype(np.float64)
p = lambda a: ctypes.c_void_p(a.ctypes.data)
out = {"n": n, "entries": int(len(I))}


def gpu_lip(T, tol, weighted, alpha, beta):
    u = np.zeros(n); sw, nl = ctypes.c_int(), ctypes.c_int()
    t0 = time.perf_counter()
    _lib.call("glb_lip_iterate_host", p(u), p(J), p(I), p(V), p(ti), p(val), T, tol, weighted, alpha, beta, n, len(I), len(ti),
              ctypes.byref(sw), ctypes.byref(nl))
    return u, sw.value, time.perf_counter() - t0


gpu_lip(5, 1e-9, 0, 0.5, 0.5)                               # warm-up (context, module load)
for name, weighted, T in (("lip_unweighted_p3", 0, 2000), ("lip_weighted_amle", 1, 300)):
    base = min(gpu_lip(0, 1e-30, weighted, 0.5, 0.5)[2] for _ in range(3))      # upload + setup + download only
    u, sw, t = min((gpu_lip(T, 1e-30, weighted, 0.5, 0.5) for _ in range(3)), key=lambda r: r[2])
    Tc = 20 if not weighted else 4
    t0 = time.perf_counter()
    if weighted:
        r, _ = c_oracle.lip_iterate_weighted(np.zeros(n), J, I, V, ti, val, Tc, 1e-30)
        chk = gpu_lip(Tc, 1e-30, 1, 0.5, 0.5)[0]
    else:
        r, _ = c_oracle.lip_iterate(np.zeros(n), J, I, V, ti, val, Tc, 1e-30, 0.5, 0.5)
        chk = gpu_lip(Tc, 1e-30, 0, 0.5, 0.5)[0]
    tc = time.perf_counter() - t0
    out[name] = {"gpu_sweeps": sw, "gpu_seconds_host_to_host": t, "gpu_setup_seconds": base,
                 "gpu_us_per_sweep": (t - base) / sw * 1e6, "cpu_oracle_ms_per_sweep_1core": None, "bit_exact_vs_oracle": bool(np.array_equal(chk, r))}
    t0 = time.perf_counter()
    if weighted:
        c_oracle.lip_iterate_weighted(np.zeros(n), J, I, V, ti, val, Tc, 1e-30)
    else:
        c_oracle.lip_iterate(np.zeros(n), J, I, V, ti, val, Tc, 1e-30, 0.5, 0.5)
    out[name]["cpu_oracle_ms_per_sweep_1core"] = (time.perf_counter() - t0) / Tc * 1e3

uu = np.ones(n); ul = np.zeros(n); uu[ti] = val; ul[ti] = val


def gpu_lp(T):
    a, b = uu.copy(), ul.copy(); sw, nl = ctypes.c_int(), ctypes.c_int()
    t0 = time.perf_counter()
    _lib.call("glb_lp_iterate_host", p(a), p(b), p(J), p(I), p(V), p(ti), p(val), 3.0, T, 1e-30, n, len(I), len(ti),
              ctypes.byref(sw), ctypes.byref(nl))
    return a, b, sw.value, time.perf_counter() - t0


gpu_lp(3)
base = min(gpu_lp(0)[3] for _ in range(3))
a, b, sw, t = min((gpu_lp(5000) for _ in range(3)), key=lambda r: r[3])
t0 = time.perf_counter(); ra, rb, _ = c_oracle.lp_iterate(uu, ul, J, I, V, ti, val, 3.0, 20, 1e-30); tc = time.perf_counter() - t0
ca, cb, _, _ = gpu_lp(20)
out["lp_jacobi_p3"] = {"gpu_sweeps": sw, "gpu_seconds_host_to_host": t, "gpu_setup_seconds": base, "gpu_us_per_sweep": (t - base) / sw * 1e6,
                       "cpu_oracle_ms_per_sweep_1core": tc / 20 * 1e3, "bit_exact_vs_oracle": bool(np.array_equal(ca, ra) and np.array_equal(cb, rb))}
print(json.dumps(out))
