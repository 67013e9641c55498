"""Smallest target for ncu: one exact kNN search at config-2 size (70 000 x 128, k+1 = 11) on device-resident features."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from graphlearning_b200 import knn_gpu
from oracle import gl_oracle as orc
n, d, k = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (70000, 128, 11)))
X, _ = orc.synthetic_blobs(n, d, c=10, seed=0)
ind, dist = knn_gpu.knnsearch_gpu(X.astype(np.float64), k)
torch.cuda.synchronize()
print("done", knn_gpu.last_stats)
