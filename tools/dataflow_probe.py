"""GPU experiment: poll statistics of the dataflow Poisson kernel (GLB_POISSON_STATS=1) on the bench graph, a tiny
graph and a chain-free case.  Not part of the product."""
import os, sys
os.environ["GLB_POISSON_STATS"] = "1"
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from graphlearning_b200 import device as gdev
from oracle import gl_oracle as orc
from test_poisson_gpu import random_knn_graph

def run(W, c, iters, tag):
    op = gdev.PoissonOperator(W, kind="dataflow")
    Db = op.pack(np.random.default_rng(0).normal(size=(W.shape[0], c)))
    u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
    for _ in range(2):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); op.iterate(Db, iters, u0, u1); e1.record(); torch.cuda.synchronize()
    print("%s: n=%d nnz=%d fill=%.3f  %.3f us/iter" % (tag, W.shape[0], W.nnz, op.fill(c), e0.elapsed_time(e1) * 1e3 / iters), flush=True)

for v in (sys.argv[1:] or ["1024,4"]):
    os.environ["GLB_POISSON_VARIANT"] = v
    print("== variant", v, flush=True)
    run(random_knn_graph(148 * 16, 4, seed=1), 10, 2000, "tiny")
    run(random_knn_graph(148 * 64, 4, seed=1), 10, 2000, "small")
    W, _ = bench.build_workload()
    run(W, 10, 1000, "bench70k")
for flags in (1, 3, 5, 7):
    os.environ["GLB_POISSON_NOPOLL"] = str(flags)
    for v in (sys.argv[1:] or ["1024,4"]):
        os.environ["GLB_POISSON_VARIANT"] = v
        print("== NOPOLL flags=%d (1 no polling, 2 no stores, 4 read u0 only) variant %s" % (flags, v), flush=True)
        W, _ = bench.build_workload()
        run(W, 10, 1000, "bench70k")
