"""GPU experiment: time the persistent Poisson kernel variants (GLB_POISSON_VARIANT=threads,unroll) on the
bench graph and on a tiny graph (barrier latency).  Prints one line per variant.  Not part of the product."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                           # noqa: E402
from graphlearning_b200 import device as gdev          # noqa: E402
from oracle import gl_oracle as orc                    # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_poisson_gpu import random_knn_graph          # noqa: E402


def time_variant(W, src, variant, iters=1000, reps=3, reorder=True, kind="auto"):
    os.environ["GLB_POISSON_VARIANT"] = variant
    op = gdev.PoissonOperator(W, reorder=reorder, kind=kind)
    Db = op.source_to_Db(src)
    u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
    best = 1e9
    for _ in range(reps + 1):
        u0.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        u, _ = op.iterate(Db, iters, u0, u1)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / iters, op.unpack(u, src.shape[1]).clone(), op.kind(src.shape[1])


def main():
    W, labels = bench.build_workload()
    ti = orc.one_per_class(labels, rate=1, seed=0)
    src = orc.poisson_source(W.shape[0], ti, labels[ti])[0]
    Wt = random_knn_graph(148 * 16, 4, seed=1)
    srct = np.random.default_rng(0).normal(size=(Wt.shape[0], 10))
    ref = None
    variants = sys.argv[1:] or ["1024,4", "1024,8", "768,8", "512,8", "512,16", "256,16"]
    for v in variants:
        us, u, pers = time_variant(W, src, v)
        usn, un, _ = time_variant(W, src, v, reorder=False)
        ust, _, _ = time_variant(Wt, srct, v, iters=2000)
        if ref is None:
            ref = un
        err = float((u - ref).abs().max() / ref.abs().max())
        print("variant %-10s kind=%s  rcm %.3f us/iter  natural %.3f us/iter  tiny graph (~barrier) %.3f us/iter  "
              "natural bit-equal=%s  rcm rel diff=%.1e" % (v, pers, us, usn, ust, bool(torch.equal(ref, un)), err), flush=True)


if __name__ == "__main__":
    main()
