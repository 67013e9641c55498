"""GPU experiment (experiment build): per-warp busy cycles of the dataflow kernel on the bench graph, written to
gpurun_out/wstats.txt (cta warp busy pairs slots) for offline analysis of the load balance."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GLB200_LIB", os.path.join(ROOT, "graphlearning_b200", "lib", "libglb200_exp.so"))
os.environ["GLB_POISSON_STATS"] = "1"
os.environ["GLB_POISSON_GATE_EVERY"] = "1"
import numpy as np, torch
import bench
from graphlearning_b200 import device as gdev
from oracle import gl_oracle as orc
W, labels = bench.build_workload()
ti = orc.one_per_class(labels, rate=1, seed=0)
op = gdev.PoissonOperator(W, kind="dataflow", reorder=True)
Db = op.source_to_Db(orc.poisson_source(W.shape[0], ti, labels[ti])[0])
u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
op.iterate(Db, 200, u0, u1)
os.environ["GLB_POISSON_WSTATS_FILE"] = os.path.join(ROOT, "gpurun_out", "wstats.txt")
u0.zero_(); op.iterate(Db, 1000, u0, u1); torch.cuda.synchronize()
print("done")
