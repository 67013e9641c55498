"""Cost of page-locked result buffers (glb_host_alloc) and of Laplace fits with / without a warm pool."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import graphlearning_b200 as gl
from graphlearning_b200 import device
from oracle import gl_oracle as orc
torch.zeros(1, device="cuda"); torch.cuda.synchronize()
held = []
for i in range(6):
    t0 = time.perf_counter(); held.append(device.pinned.empty((70000, 10))); t = time.perf_counter() - t0
    print("pinned.empty #%d (5.6 MB, new allocation): %.2f ms" % (i, 1e3 * t), flush=True)
del held
t0 = time.perf_counter(); a = device.pinned.empty((70000, 10)); print("from the pool: %.3f ms" % (1e3 * (time.perf_counter() - t0)))
for i in range(3):
    t0 = time.perf_counter(); b = np.empty((70000, 10)); b[:] = 0; print("np.empty + touch: %.2f ms" % (1e3 * (time.perf_counter() - t0)))
X, labels = orc.synthetic_blobs(70000, 8, c=10, seed=0)
W = gl.weightmatrix.knn(X.astype(np.float64), 10)
t5 = orc.one_per_class(labels, rate=5, seed=0)
m = gl.ssl.laplace(W)
for i in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter(); m.fit(t5, labels[t5]); t = time.perf_counter() - t0
    print("laplace fit #%d: %.2f ms (CG %.2f ms)" % (i, 1e3 * t, m.cg_info["device_ms"]), flush=True)
os.environ["GLB_TIMING"] = "1"
m.fit(t5, labels[t5])
