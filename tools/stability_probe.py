"""GPU experiment: run-to-run stability of the dataflow kernel (is the low-polling regime always reached?)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import graphlearning_b200 as gl
from graphlearning_b200 import device as gdev
from oracle import gl_oracle as orc
W, labels = bench.build_workload()
ti = orc.one_per_class(labels, rate=1, seed=0)
src = orc.poisson_source(W.shape[0], ti, labels[ti])[0]
for kind in ("dataflow", "barrier"):
    op = gdev.PoissonOperator(W, kind=kind)
    Db = op.source_to_Db(src)
    u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for mode in ("noflush", "flush"):
        ts = []
        for i in range(40 if kind == 'dataflow' else 8):
            u0.zero_()
            if mode == "flush": flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); op.iterate(Db, 1000, u0, u1); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(kind, mode, "ms per 1000 iterations:", " ".join("%.2f" % t for t in ts), flush=True)
model = gl.ssl.poisson(W, solver="gradient_descent", min_iter=1000, max_iter=1000)
ts = []
for i in range(40):
    torch.cuda.synchronize(); t0 = time.perf_counter(); model.fit(ti, labels[ti]); ts.append((time.perf_counter() - t0) * 1e3)
print("e2e fit ms:", " ".join("%.2f" % t for t in ts), flush=True)
