"""World-size-2 test of the row-partitioned iterate's host logic over gloo on CPU: partition,
slab extraction with global columns, padded layout, all-gather protocol.  The local product is scipy here (the
oracle's arithmetic, test only); on the GPU box the same protocol drives poisson_step_kernel over NCCL
(tests/test_distributed_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy import sparse

from graphlearning_b200 import distributed as gd
from oracle import gl_oracle as orc


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _graph(n=900, k=7, seed=0):
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), k); cols = rng.integers(0, n, n * k)
    W = sparse.coo_matrix((np.exp(-4 * rng.random(n * k)), (rows, cols)), shape=(n, n)).tocsr()
    W = sparse.csr_matrix((W + W.T) / 2); W.setdiag(0); W.eliminate_zeros()
    return W


def _worker(rank, world, port, T, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W = _graph()
    n = W.shape[0]
    s = orc.poisson_gd_setup(W, np.array([0, 1, 2]), np.array([0, 1, 2]))
    P, Db = s["P"], s["Db"]
    c = Db.shape[1]
    bounds = gd.partition_rows(P.indptr, world)
    proto = gd.PartitionedIterate(bounds, rank, world, c,
                                  make_buffer=lambda r, l: torch.zeros((r, l), dtype=torch.float64),
                                  all_gather=lambda dst, slab: dist.all_gather_into_tensor(dst, slab))
    rp, col, val = gd.row_slab(P, proto.r0, proto.r1)
    pos = proto.padded_index()
    Ploc = sparse.csr_matrix((val, pos[col], rp), shape=(proto.r1 - proto.r0, world * proto.rows_pad))

    def local_step(u_full, out_slab):
        out_slab.zero_()
        out_slab[: proto.r1 - proto.r0] = torch.from_numpy(Db[proto.r0:proto.r1] + Ploc @ u_full.numpy())

    out = proto.run(local_step, T)
    np.save(os.path.join(out_dir, "u%d.npy" % rank), out.numpy()[pos])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_iterate_matches_single_process(tmp_path, world):
    T = 9
    mp.spawn(_worker, args=(world, _free_port(), T, str(tmp_path)), nprocs=world, join=True)
    W = _graph()
    s = orc.poisson_gd_setup(W, np.array([0, 1, 2]), np.array([0, 1, 2]))
    u = np.zeros_like(s["Db"])
    for _ in range(T):
        u = s["Db"] + s["P"] * u
    for r in range(world):
        got = np.load(tmp_path / ("u%d.npy" % r))
        assert np.array_equal(got, u)             # same per-row arithmetic (scipy csr_matvecs), every rank has everything


def test_partition_is_balanced_and_covers_all_rows():
    W = _graph(5000, 9, seed=3)
    for world in (1, 2, 4, 8):
        b = gd.partition_rows(W.indptr, world)
        assert b[0] == 0 and b[-1] == 5000 and np.all(np.diff(b) >= 0) and len(b) == world + 1
        nz = np.diff(W.indptr[b])
        assert nz.max() <= 1.05 * W.nnz / world + 50
    # degenerate: more ranks than rows
    b = gd.partition_rows(np.array([0, 3, 5]), 4)
    assert b[0] == 0 and b[-1] == 2 and np.all(np.diff(b) >= 0)
    rp, col, val = gd.row_slab(sparse.csr_matrix(np.eye(3)), 1, 1)
    assert list(rp) == [0, 0] and len(col) == 0
