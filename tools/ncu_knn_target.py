"""ncu target: the fused kNN search.  usage: python tools/ncu_knn_target.py [n] [d] [k]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from graphlearning_b200 import knn_gpu
from oracle import gl_oracle as orc
n = int(sys.argv[1]) if len(sys.argv) > 1 else 37888
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
k = int(sys.argv[3]) if len(sys.argv) > 3 else 11
X, _ = orc.synthetic_blobs(n, d, c=10, seed=0)
ind, dist = knn_gpu.knnsearch_gpu(X.astype(np.float64), k)
print("done", knn_gpu.last_stats)
