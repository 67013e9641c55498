"""Row-partitioned Poisson iterate on the GPU: world size 1 in-process; world size 2 over NVLink / NCCL when the box has
two GPUs (skipped otherwise).  The halo-put kernel (csrc/slab.cu), the all-gather baseline and the single-GPU step kernel
must agree bitwise (same per-row arithmetic in the same order)."""
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


def _hub_graph(n, seed):
    from scipy import sparse
    rng = np.random.default_rng(seed)
    W = random_knn_graph(n, 9, seed=seed).tolil()
    for hub_node, deg in ((3, 700), (n - 2, 150)):                 # rows longer than a slice: dealt over a whole warp
        hub = rng.choice(n, deg, replace=False)
        hub = hub[hub != hub_node]
        W[hub_node, hub] = 0.5; W[hub, hub_node] = 0.5
    W = sparse.csr_matrix(W); W.eliminate_zeros()
    return W


@pytest.mark.parametrize("reorder", [False, True])
@pytest.mark.parametrize("c", [10, 3, 40])
def test_slab_kernel_world_one(reorder, c):
    """glb_slab_* on one GPU (no peers): against the plain-C oracle and bitwise against the step kernel on the same fp32 P
    (the all-gather baseline with one rank: same host-built slab, poisson_step_kernel)."""
    from graphlearning_b200 import device as gdev, distributed as gd
    n = 6000
    W = _hub_graph(n, seed=4)
    src = np.random.default_rng(0).normal(size=(n, c)) * (np.random.default_rng(1).random((n, 1)) < 0.01)
    pp = gd.PartitionedPoisson(W, rank=0, world=1, reorder=reorder, c=c)
    u = pp.iterate(src, 17)
    u_again = pp.iterate(src, 17)
    u4 = pp.iterate(src, 4)
    pp.close()
    assert np.array_equal(u, u_again)
    s = orc.poisson_gd_setup(W, np.array([0]), np.array([0]))
    Db = (1.0 / (W * np.ones(n)))[:, None] * src
    assert rel_err(u, c_oracle.poisson_iterate(s["P"], Db, 17)) <= 1e-5
    assert rel_err(u4, c_oracle.poisson_iterate(s["P"], Db, 4)) <= 1e-5
    # without rows longer than a slice (those are summed by a whole warp, in another order) the arithmetic is the step kernel's
    W2 = random_knn_graph(n, 9, seed=5)
    pp = gd.PartitionedPoisson(W2, rank=0, world=1, reorder=reorder, c=c)
    u2 = pp.iterate(src, 17)
    pp.close()
    ref = gd.AllGatherPoisson(W2, rank=0, world=1, reorder=reorder).iterate(src, 17)
    assert np.array_equal(u2, ref), float(np.abs(u2 - ref).max())


def test_slab_kernel_big_tiles():
    """A slab with many tiles per CTA gets tiles of 32 slices instead of 16 (a million rows on 148 SMs): same arithmetic,
    bitwise the step kernel's result, and the choice is really taken."""
    from graphlearning_b200 import distributed as gd
    n, c = 1000000, 10
    W = random_knn_graph(n, 5, seed=11)
    src = np.zeros((n, c))
    lab = np.random.default_rng(2).choice(n, 200, replace=False)
    src[lab] = np.random.default_rng(3).normal(size=(200, c))
    pp = gd.PartitionedPoisson(W, rank=0, world=1, reorder=False, c=c)
    assert pp.tile_slices == 32
    u = pp.iterate(src, 6)
    pp.close()
    ref = gd.AllGatherPoisson(W, rank=0, world=1, reorder=False).iterate(src, 6)
    assert np.array_equal(u, ref), float(np.abs(u - ref).max())
    small = gd.PartitionedPoisson(random_knn_graph(6000, 9, seed=5), rank=0, world=1, reorder=False, c=c)
    assert small.tile_slices == 16
    small.close()


def test_allgather_baseline_world_one():
    from graphlearning_b200 import device as gdev, distributed as gd
    W = random_knn_graph(6000, 9, seed=4)
    src = np.random.default_rng(0).normal(size=(6000, 10)) * (np.random.default_rng(1).random((6000, 1)) < 0.01)
    pp = gd.AllGatherPoisson(W, rank=0, world=1)
    u = pp.iterate(src, 17)
    op = gdev.PoissonOperator(W, kind="step")
    ref = op.unpack(op.iterate(op.source_to_Db(src), 17)[0], 10).cpu().numpy()
    # P is rounded to fp32 from D^-1 * w on the host here (as the reference forms it, ssl.py:634-635) and from w / d on the
    # device there: the last bit of a few entries differs
    assert rel_err(u, ref) <= 1e-6


@pytest.mark.parametrize("reorder", [1, 0])
def test_world_size_two_over_nvlink(tmp_path, reorder):
    """two ranks: halo rows put into the peer's label matrix by the kernel == all-gather baseline == one GPU, bitwise"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from graphlearning_b200 import distributed as gd
    out = tmp_path / "res.npz"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29671", os.path.join(ROOT, "tools", "bench_cfg5.py"), "--check", str(out), "--size", "20000", "--iters", "15",
           "--reorder", str(reorder)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    z = np.load(out)
    assert np.array_equal(z["u_put"], z["u_allgather"])
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_cfg5
    W = bench_cfg5.build_graph(20000)
    pp = gd.PartitionedPoisson(W, rank=0, world=1, reorder=bool(reorder), c=10)
    single = pp.iterate(z["src"], 15)
    pp.close()
    assert np.array_equal(z["u_put"], single)
