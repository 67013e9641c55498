"""GPU experiment: where the time of the FIRST fit on a graph goes (GLB_TIMING=1 phases + Python-side wall clock)."""
import gc, os, sys, time
os.environ["GLB_TIMING"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import graphlearning_b200 as gl
from oracle import gl_oracle as orc
W, labels = bench.build_workload()
ti = orc.one_per_class(labels, rate=1, seed=0)
torch.zeros(1, device="cuda"); torch.cuda.synchronize()
m = None
for rep in range(3):
    m = None
    gc.collect(); torch.cuda.synchronize()               # the previous graph's device state is freed BEFORE the clock starts
    Wc = W.copy()
    t0 = time.perf_counter()
    m = gl.ssl.poisson(Wc, solver="gradient_descent", min_iter=200, max_iter=200)
    t1 = time.perf_counter()
    m.fit(ti, labels[ti])
    t2 = time.perf_counter()
    m.fit(ti, labels[ti])
    t3 = time.perf_counter()
    print("rep %d: construct %.1f ms, first fit %.1f ms, second fit %.1f ms" % (rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
