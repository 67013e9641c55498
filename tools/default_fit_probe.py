"""Where the time of a default fit (min_iter=50, stopping rule on) goes: the mixing kernel alone, the whole fit, repeated."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import graphlearning_b200 as gl
from graphlearning_b200 import device
from oracle import gl_oracle as orc
W, labels = bench.build_workload()
ti = orc.one_per_class(labels, rate=1, seed=0)
op = device.PoissonOperator(W)
for lo, hi in ((50, 1000), (64, 64), (200, 200), (1, 1)):
    op.mixing_T(ti, lo, hi); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        T = op.mixing_T(ti, lo, hi)
    torch.cuda.synchronize()
    print("mixing_T(%d, %d) = %d: %.3f ms per call" % (lo, hi, T, (time.perf_counter() - t0) / 5 * 1e3), flush=True)
for kw in (dict(), dict(min_iter=50, max_iter=50), dict(min_iter=1000, max_iter=1000)):
    m = gl.ssl.poisson(W, solver="gradient_descent", **kw)
    m.fit(ti, labels[ti]); m.fit(ti, labels[ti])
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        u = m.fit(ti, labels[ti])
    torch.cuda.synchronize()
    print("fit %s: T = %d, %.3f ms per fit" % (kw, m.iterations, (time.perf_counter() - t0) / 5 * 1e3), flush=True)
os.environ["GLB_TIMING"] = "1"
m = gl.ssl.poisson(W, solver="gradient_descent")
m.fit(ti, labels[ti])
