"""Laplace-learning fits whose CG kernels are to be listed by ncu (launch list) or timed: the 70k d=8 blob graph of the
bench (95 iterations) and the config-3 graph (60 000 x 512, k = 20; hub rows of thousands of nonzeros).
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cg_launches.csv python tools/cg_probe.py
  python tools/cg_probe.py time     # wall/device times only
"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
import graphlearning_b200 as gl
from oracle import gl_oracle as orc

which = sys.argv[1] if len(sys.argv) > 1 else "both"
X, labels = orc.synthetic_blobs(70000, 8, c=10, seed=0)
W = gl.weightmatrix.knn(X.astype(np.float64), 10)
t5 = orc.one_per_class(labels, rate=5, seed=0)
m = gl.ssl.laplace(W)
m.fit(t5, labels[t5])
torch.cuda.synchronize(); t0 = time.perf_counter(); m.fit(t5, labels[t5]); t = time.perf_counter() - t0
print("70k: fit %.2f ms, %d CG iterations, device %.3f ms = %.1f us/iteration, launches %d" %
      (1e3 * t, m.iterations, m.cg_info["device_ms"], 1e3 * m.cg_info["device_ms"] / m.iterations, m.gpu_launches), flush=True)
X3, lab3 = orc.synthetic_blobs(60000, 512, c=10, seed=0)
W3 = gl.weightmatrix.knn(X3.astype(np.float64), 20)
rl = np.diff(W3.indptr)
print("cfg3 graph: nnz %d, max row %d, rows > 64: %d, rows > 768: %d" % (W3.nnz, rl.max(), (rl > 64).sum(), (rl > 768).sum()))
t3 = orc.one_per_class(lab3, rate=5, seed=0)
m3 = gl.ssl.laplace(W3)
m3.fit(t3, lab3[t3])
for tol in (1e-5, 1e-12):
    m3.tol = tol
    torch.cuda.synchronize(); t0 = time.perf_counter(); m3.fit(t3, lab3[t3]); t = time.perf_counter() - t0
    print("cfg3 tol %g: fit %.2f ms, %d CG iterations, device %.3f ms = %.1f us/iteration" %
          (tol, 1e3 * t, m3.iterations, m3.cg_info["device_ms"], 1e3 * m3.cg_info["device_ms"] / m3.iterations), flush=True)
