"""Row-partitioned Poisson iterate on the GPU: world size 1 in-process; world size 2 over NCCL when the box has
two GPUs (skipped otherwise).  Results must be bitwise those of the single-GPU step kernel."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_err
from oracle import c_oracle
from oracle import gl_oracle as orc
from test_poisson_gpu import random_knn_graph

pytestmark = pytest.mark.gpu


def test_world_size_one_equals_step_kernel():
    from graphlearning_b200 import device as gdev, distributed as gd
    W = random_knn_graph(6000, 9, seed=4)
    src = np.random.default_rng(0).normal(size=(6000, 10)) * (np.random.default_rng(1).random((6000, 1)) < 0.01)
    pp = gd.PartitionedPoisson(W, rank=0, world=1)
    u = pp.iterate(src, 17)
    op = gdev.PoissonOperator(W, kind="step")
    ref = op.unpack(op.iterate(op.source_to_Db(src), 17)[0], 10).cpu().numpy()
    assert np.array_equal(u, ref), float(np.abs(u - ref).max())
    s = orc.poisson_gd_setup(W, np.array([0]), np.array([0]))
    oracle = c_oracle.poisson_iterate(s["P"], (1.0 / (W * np.ones(6000)))[:, None] * src, 17)
    assert rel_err(u, oracle) <= 1e-5


def test_world_size_two_over_nccl(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = tmp_path / "res.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29671", os.path.join(ROOT, "tools", "bench_cfg5.py"), "--check", str(out), "--size", "20000", "--iters", "15"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    z = np.load(out)
    assert np.array_equal(z["u_partitioned"], z["u_single"])
