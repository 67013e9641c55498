"""ncu target: the row-slab step kernel on the config-5 graph (one GPU).  usage: python tools/ncu_slab_target.py [n] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import bench_cfg5
from graphlearning_b200 import distributed as gd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
W = bench_cfg5.build_graph(n)
pp = gd.PartitionedPoisson(W, rank=0, world=1, reorder=True, c=10)
print("fill", pp._lib.load().glb_slab_fill(pp._slab), "ms/iter", pp.timed_iterations(iters) / iters)
pp.close()
