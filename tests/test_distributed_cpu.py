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


# ---- halo exchange (PartitionedPoisson's host logic) ------------------------------------------------------------------------
def _local_graph(n=1500, k=6, seed=1):
    """a graph with locality (points on a line, neighbours by rank distance) plus a few long-range edges and one hub"""
    rng = np.random.default_rng(seed)
    rows = np.repeat(np.arange(n), k)
    cols = np.clip(rows + rng.integers(-25, 26, n * k), 0, n - 1)
    far = rng.random(n * k) < 0.01
    cols[far] = rng.integers(0, n, far.sum())
    W = sparse.coo_matrix((0.1 + rng.random(n * k), (rows, cols)), shape=(n, n)).tocsr()
    W = sparse.csr_matrix((W + W.T) / 2)
    hub = rng.choice(n, 300, replace=False)
    H = sparse.coo_matrix((np.full(300, 0.3), (np.full(300, 7), hub)), shape=(n, n)).tocsr()
    W = sparse.csr_matrix(W + H + H.T); W.setdiag(0); W.eliminate_zeros()
    return W


def test_poisson_slab_equals_rows_of_P():
    """poisson_slab builds rows [r0, r1) of P = D^-1 W^T in a relabelled numbering without forming P."""
    W = _local_graph(directed := 700, 5, seed=4)
    W = sparse.csr_matrix(W + sparse.random(700, 700, 0.002, random_state=1, format="csr"))       # not symmetric any more
    W.setdiag(0); W.eliminate_zeros()
    rng = np.random.default_rng(0)
    perm = rng.permutation(700)
    s = orc.poisson_gd_setup(W, np.array([0]), np.array([0]))
    Pp = sparse.csr_matrix(s["P"][perm][:, perm]); Pp.sort_indices(); Pp.eliminate_zeros()
    for r0, r1 in ((0, 700), (100, 333), (650, 700), (5, 5)):
        rp, col, val, deg = gd.poisson_slab(W, perm, r0, r1)
        a, b = Pp.indptr[r0], Pp.indptr[r1]
        assert np.array_equal(rp, Pp.indptr[r0:r1 + 1] - a)
        assert np.array_equal(col, Pp.indices[a:b])
        assert np.array_equal(val, Pp.data[a:b].astype(np.float32))
        assert np.array_equal(deg, np.asarray(W.sum(axis=1)).ravel()[perm[r0:r1]])


def _halo_worker(rank, world, port, T, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    W = _local_graph()
    n = W.shape[0]
    perm = np.random.default_rng(5).permutation(n) if rank >= 0 else None      # any relabelling: the protocol must not care
    lens = np.bincount(W.indices, minlength=n)[perm]
    bounds = gd.partition_rows(np.concatenate(([0], np.cumsum(lens))), world)
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    rp, colg, val, deg = gd.poisson_slab(W, perm, r0, r1)
    part = gd.HaloPartition(rp, colg, val, bounds, rank, allgather)
    c = 3
    src = np.random.default_rng(9).normal(size=(n, c)) * (np.random.default_rng(8).random((n, 1)) < 0.05)
    Db = (1.0 / deg)[:, None] * src[perm[r0:r1]]
    u = np.zeros((part.rows_total[rank], c))
    assert part.rows_total[rank] == part.m + len(part.halo) + 1
    # invariants of the structure
    assert np.all(part.boundary[np.diff(part.send_ptr) > 0] == 1)
    assert part.col.max(initial=0) < part.m + len(part.halo)
    for t in range(T):
        new = part.step_numpy(Db, u)
        # interior rows must not depend on anything a peer delivers
        P = sparse.csr_matrix((part.val.astype(np.float64), part.col, part.rp), shape=(part.m, len(u)))
        assert P[part.boundary == 0][:, part.m:].nnz == 0
        boxes = allgather(part.puts(new))                             # every rank's outgoing puts
        u[: part.m] = new
        for sender in range(world):
            if rank in boxes[sender]:
                assert (part.neighbours >> sender) & 1
                dst, vals = boxes[sender][rank]
                assert dst.min() >= part.m and dst.max() < part.m + len(part.halo)
                u[dst] = vals
    np.savez(os.path.join(out_dir, "h%d.npz" % rank), rows=perm[r0:r1], u=u[: part.m], halo=perm[part.halo], uh=u[part.m:part.m + len(part.halo)])
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_matches_single_process(tmp_path, world):
    T = 7
    mp.spawn(_halo_worker, args=(world, _free_port(), T, str(tmp_path)), nprocs=world, join=True)
    W = _local_graph()
    n = W.shape[0]
    s = orc.poisson_gd_setup(W, np.array([0]), np.array([0]))
    P32 = sparse.csr_matrix((s["P"].data.astype(np.float32).astype(np.float64), s["P"].indices, s["P"].indptr), shape=(n, n))
    src = np.random.default_rng(9).normal(size=(n, 3)) * (np.random.default_rng(8).random((n, 1)) < 0.05)
    Db = (1.0 / np.asarray(W.sum(axis=1)).ravel())[:, None] * src
    u = np.zeros_like(Db)
    for _ in range(T):
        u = Db + P32 @ u
    seen = np.zeros(n, dtype=bool)
    for r in range(world):
        z = np.load(tmp_path / ("h%d.npz" % r))
        assert np.allclose(z["u"], u[z["rows"]], rtol=1e-12, atol=1e-15)       # row sums in a relabelled column order
        assert np.allclose(z["uh"], u[z["halo"]], rtol=1e-12, atol=1e-15)      # the halo holds the owners' latest rows
        seen[z["rows"]] = True
    assert seen.all()
