"""Smallest possible target for ncu: build the bench graph's operator, run the iterate a few times.
usage: python tools/ncu_target.py [iters] [calls] [reorder]     (GLB200_LIB / GLB_POISSON_* select the experiment build's switches)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from graphlearning_b200 import device as gdev
from oracle import gl_oracle as orc
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 100
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
reorder = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
W, labels = bench.build_workload()
ti = orc.one_per_class(labels, rate=1, seed=0)
op = gdev.PoissonOperator(W, kind=os.environ.get("GLB_KIND", "auto"), reorder=reorder)
Db = op.source_to_Db(orc.poisson_source(W.shape[0], ti, labels[ti])[0])
u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
for _ in range(calls):
    u0.zero_()
    op.iterate(Db, iters, u0, u1)
torch.cuda.synchronize()
print("done", op.kind(10), "gate", op.gate(10))
