"""GPU experiment: A/B of the dataflow Poisson kernel generations and CTA shapes on the bench graph, same process, same
box.  GLB_POISSON_DF=1 is the first-generation kernel.  Prints one line per configuration.  Not part of the product."""
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
configs = [("1", "512,16", "0"), ("2", "512,16", "0"), ("2", "512,8", "0"), ("2", "1024,8", "0"), ("2", "768,8", "0"), ("2", "256,16", "0"),
           ("1", "512,16", "1"), ("2", "512,16", "1"), ("2", "1024,8", "1"), ("2", "512,16", "3"), ("1", "512,16", "0"), ("2", "512,16", "0")]
if len(sys.argv) > 1 and sys.argv[1] == "short":
    configs = [("1", "512,16", "0"), ("2", "512,8", "0"), ("2", "512,8", "1"), ("2", "512,8", "3"), ("1", "512,16", "0"), ("2", "512,8", "0")]
for df, variant, nopoll in configs:
    os.environ["GLB_POISSON_DF"] = df
    os.environ["GLB_POISSON_VARIANT"] = variant
    if nopoll != "0":
        os.environ["GLB_POISSON_NOPOLL"] = nopoll
    else:
        os.environ.pop("GLB_POISSON_NOPOLL", None)
    op = gdev.PoissonOperator(W, kind="dataflow")
    Db = op.source_to_Db(src)
    u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
    times = []
    for _ in range(5):
        u0.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); u, _ = op.iterate(Db, 1000, u0, u1); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    res = op.unpack(u, 10).clone()
    if ref is None:
        ref = res
    err = float((res - ref).abs().max() / ref.abs().max())
    print("DF=%s variant=%-8s nopoll=%s  gate=%d fill=%.3f  us/iter: best %.3f median %.3f  rel diff vs first=%.1e" % (
        df, variant, nopoll, op.gate(10), op.fill(10), min(times), float(np.median(times)), err), flush=True)
    del op
