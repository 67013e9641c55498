"""GPU experiment: dataflow kernel with the first gather attempt served through L1 (GLB_POISSON_NOPOLL=8), natural vs RCM
ordering, on the bench graph.  Results must be bit-identical to the default path.  Not part of the product."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from graphlearning_b200 import device as gdev
from oracle import gl_oracle as orc
W, labels = bench.build_workload()
ti = orc.one_per_class(labels, rate=1, seed=0)
src = orc.poisson_source(W.shape[0], ti, labels[ti])[0]
ref = None
for reorder in (False, True):
    os.environ.pop("GLB_POISSON_NOPOLL", None)
    op = gdev.PoissonOperator(W, kind="dataflow", reorder=reorder)
    Db = op.source_to_Db(src)
    u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
    for flag in ("0", "8", "0", "8"):
        if flag == "0":
            os.environ.pop("GLB_POISSON_NOPOLL", None)
        else:
            os.environ["GLB_POISSON_NOPOLL"] = flag
        times = []
        for _ in range(4):
            u0.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); u, _ = op.iterate(Db, 1000, u0, u1); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        res = op.unpack(u, 10).clone()
        if ref is None:
            ref = res
        print("reorder=%s l1_first=%s gate=%d: best %.3f median %.3f us/iter, max rel diff vs first %.1e" % (
            reorder, flag, op.gate(10), min(times), float(np.median(times)), float((res - ref).abs().max() / ref.abs().max())), flush=True)
