"""GPU experiment: A/B of the dataflow Poisson kernel variants on the bench graph (and optionally the 128-d hub graph),
same process, same box.  Needs the -DGLB_EXPERIMENT library (python -m graphlearning_b200.build --exp), whose plan
creation reads the switches below.  Prints one line per configuration.  Not part of the product.

    GLB200_LIB=graphlearning_b200/lib/libglb200_exp.so python tools/df_ab.py [hub] [stats]

switches: GLB_POISSON_L1 (first gather attempt through L1),
GLB_POISSON_THREADS (CTA size of the pipelined kernel), reorder (RCM locality ordering of the operator)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("GLB200_LIB", os.path.join(ROOT, "graphlearning_b200", "lib", "libglb200_exp.so"))
import numpy as np, torch
import bench
from graphlearning_b200 import device as gdev
from oracle import gl_oracle as orc, c_oracle

hub = "hub" in sys.argv
if "stats" in sys.argv:
    os.environ["GLB_POISSON_STATS"] = "1"
if hub:
    from graphlearning_b200 import knn_gpu
    X, labels = orc.synthetic_blobs(70000, 128, c=10, seed=0)
    ind, dist_ = knn_gpu.knnsearch_gpu(X.astype(np.float64), 11)
    W = orc.knn_weights(ind, dist_, 10)
else:
    W, labels = bench.build_workload()
n = W.shape[0]
ti = orc.one_per_class(labels, rate=1, seed=0)
src = orc.poisson_source(n, ti, labels[ti])[0]
s = orc.poisson_gd_setup(W, ti, labels[ti])
ref50 = c_oracle.poisson_iterate(s["P"], np.asarray(s["Db"]), 50)
b_iter = bench.algorithmic_bytes(n, W.nnz, 10)
print("graph: n=%d nnz=%d max row %d; bytes/iteration %d" % (n, W.nnz, int(np.diff(W.indptr).max()), b_iter), flush=True)

def run(label, env, reorder):
    global first
    for k in ("GLB_POISSON_L1", "GLB_POISSON_THREADS", "GLB_POISSON_SLEEP", "GLB_POISSON_GATE_EVERY", "GLB_POISSON_FREE"):
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        op = gdev.PoissonOperator(W, kind="dataflow", reorder=bool(reorder))
        Db = op.source_to_Db(src)
        u50 = op.unpack(op.iterate(Db, 50)[0], 10).cpu().numpy()
        err = float(np.abs(u50 - ref50).max() / np.abs(ref50).max())
        u0 = torch.zeros_like(Db); u1 = torch.zeros_like(Db)
        times = []
        for _ in range(5):
            u0.zero_()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); u, _ = op.iterate(Db, 1000, u0, u1); e1.record(); torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        res = op.unpack(u, 10).clone()
        if first is None:
            first = res
        d = float((res - first).abs().max() / first.abs().max())
        best = min(times)
        print("%-44s reorder=%d gate=%-2d fill=%.3f  us/iter best %.3f median %.3f worst %.3f  frac %.3f  err@50 %.1e  diff@1000 %.1e" % (
            label, reorder, op.gate(10), op.fill(10), best, float(np.median(times)), max(times),
            b_iter * 1000 / (best * 1e-3) / 1e9 / 6451.8, err, d), flush=True)
    except Exception as e:
        print("%s reorder=%d FAILED: %r" % (label, reorder, e), flush=True)


first = None
if "free" in sys.argv:
    # ceiling probes: no synchronisation at all (free=1), and no stores either (free=3); results are wrong by design
    for l1, reorder in ((1, 1), (1, 0), (0, 0)):
        for free in (0, 1, 3):
            run("l1=%d free=%d" % (l1, free), {"GLB_POISSON_L1": l1, "GLB_POISSON_FREE": free, "GLB_POISSON_GATE_EVERY": 1 if free == 0 else 0}, reorder)
elif "sleep" in sys.argv:
    for l1, reorder in ((1, 1), (0, 0)):
        for sleep in (0, 100, 400, 1500):
            for gate in (0, 32, 4, 1):
                run("l1=%d sleep=%d gate_forced=%d" % (l1, sleep, gate), {"GLB_POISSON_L1": l1, "GLB_POISSON_SLEEP": sleep, "GLB_POISSON_GATE_EVERY": gate}, reorder)
else:
    for l1, threads, reorder in ((1, 512, 1), (1, 512, 0), (0, 512, 0), (0, 512, 1), (1, 768, 1), (1, 1024, 1)):
        run("l1=%d threads=%d" % (l1, threads), {"GLB_POISSON_L1": l1, "GLB_POISSON_THREADS": threads}, reorder)
