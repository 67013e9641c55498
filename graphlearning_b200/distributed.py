"""Row-partitioned Poisson iterate across the GPUs of one node (BASELINE config 5).

One process per GPU (torchrun).  Rank g owns a contiguous block of rows of P = D^-1 W^T (CSR slab with GLOBAL
column indices) and the same rows of Db; every rank keeps two full-size label matrices.  One iteration is

    local rows of u_{t+1}  <-  Db_local + P_local u_t          (poisson_step_kernel through the C-ABI)
    all ranks              <-  all-gather of the row slabs     (NCCL over NVLink; gloo in the CPU tests)

which is the reference loop `u = Db + P*u` (graphlearning/ssl.py:668) with the rows dealt out; the results are
bitwise those of the single-GPU step kernel (same per-row arithmetic, no cross-rank reduction).

The only collective is the all-gather: every rank may need any row of u_t (kNN columns are arbitrary).  Row blocks
are balanced by nonzeros and padded to equal length so that `all_gather_into_tensor` applies.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse


def partition_rows(indptr, world):
    """Contiguous row blocks with (almost) equal nonzero counts: bounds[g]..bounds[g+1] for rank g."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n = len(indptr) - 1
    cost = indptr + 2 * np.arange(n + 1)                       # nonzeros + a constant per row
    targets = cost[-1] * np.arange(1, world) / world
    inner = np.searchsorted(cost, targets, side="left")
    bounds = np.concatenate(([0], np.minimum(inner, n), [n])).astype(np.int64)
    return np.maximum.accumulate(bounds)


def row_slab(P, r0, r1):
    """CSR slab of rows [r0, r1) with global column indices: (rowptr int32, col int32, val)."""
    P = sparse.csr_matrix(P)
    a, b = P.indptr[r0], P.indptr[r1]
    rp = (P.indptr[r0:r1 + 1] - a).astype(np.int32)
    if r1 == r0:
        rp = np.zeros(2, dtype=np.int32)                   # a rank without rows still gets a well-formed (1-row, empty) slab
    return rp, P.indices[a:b].astype(np.int32), P.data[a:b]


class PartitionedIterate:
    """The exchange protocol, independent of where the local product runs.

    local_step(u_full, out_slab): writes rows [r0, r1) of Db + P u_full into out_slab[: r1 - r0].
    `u` buffers are (world * rows_pad, ld) tensors: rank g's rows live at [g * rows_pad, g * rows_pad + len_g).
    """

    def __init__(self, bounds, rank, world, ld, make_buffer, all_gather):
        self.bounds = np.asarray(bounds, dtype=np.int64)
        self.rank, self.world, self.ld = rank, world, ld
        self.rows_pad = int(np.max(np.diff(self.bounds))) if world > 0 else 0
        self.r0, self.r1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.all_gather = all_gather
        self.u = [make_buffer(world * self.rows_pad, ld), make_buffer(world * self.rows_pad, ld)]
        self.slab = make_buffer(self.rows_pad, ld)

    def padded_index(self):
        """padded row position of every global row: global row i of rank g sits at g * rows_pad + (i - bounds[g])."""
        n = int(self.bounds[-1])
        owner = np.searchsorted(self.bounds, np.arange(n), side="right") - 1
        return owner * self.rows_pad + (np.arange(n) - self.bounds[owner])

    def run(self, local_step, T):
        """T iterations from u[0]; returns the buffer that holds the result."""
        for t in range(T):
            src, dst = self.u[t & 1], self.u[(t + 1) & 1]
            local_step(src, self.slab)
            self.all_gather(dst, self.slab)
        return self.u[T & 1]


class PartitionedPoisson:
    """Device side: row slab of P on this rank's GPU, step kernel + NCCL all-gather.  Needs torch.distributed
    initialised with the nccl backend and one GPU per rank."""

    def __init__(self, W, rank=None, world=None):
        import torch
        import torch.distributed as dist
        from . import device as gdev
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        W = sparse.csr_matrix(W)
        self.n = W.shape[0]
        # P = D^-1 W^T in fp64 on the host exactly as ssl.py:634-635, rounded to fp32 once (as the single-GPU path)
        W0 = W - sparse.spdiags(W.diagonal(), 0, self.n, self.n)
        deg = np.asarray(W0.sum(axis=1)).ravel()
        P = sparse.csr_matrix(sparse.spdiags(1.0 / deg, 0, self.n, self.n) * W0.T)
        P.sort_indices()                       # the single-GPU path sums every row in ascending column order
        self.deg = deg
        self.bounds = partition_rows(P.indptr, self.world)
        r0, r1 = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        rp, col, val = row_slab(P, r0, r1)
        # columns point into the PADDED layout of the full label matrix
        self._proto = None
        self._slab_host = (rp, col, val.astype(np.float32))
        self.nnz_local = len(col)
        self._torch, self._dist, self._gdev = torch, dist, gdev
        self._plans = {}

    def _setup(self, c):
        torch, dist, gdev = self._torch, self._dist, self._gdev
        from . import _lib
        import ctypes
        ld = _lib.padded_ld(c)
        dev = torch.device("cuda", torch.cuda.current_device())
        proto = PartitionedIterate(
            self.bounds, self.rank, self.world, ld,
            make_buffer=lambda r, l: torch.zeros((r, l), dtype=torch.float32, device=dev),
            all_gather=lambda dst, slab: dist.all_gather_into_tensor(dst, slab) if self.world > 1 else dst[: slab.shape[0]].copy_(slab))
        rp, col, val = self._slab_host
        pos = proto.padded_index()
        self.rowptr = torch.from_numpy(rp).to(dev)
        self.col = torch.from_numpy(pos[col].astype(np.int32)).to(dev)
        self.val = torch.from_numpy(val).to(dev)
        n_local = proto.r1 - proto.r0
        h = ctypes.c_void_p()
        _lib.call("glb_poisson_plan_create", ctypes.byref(h), gdev.ptr(self.rowptr), gdev.ptr(self.col), gdev.ptr(self.val),
                  max(n_local, 1), self.nnz_local, c, 0, gdev.cur_stream())            # kind 0 = one launch per iteration
        self._plans[c] = (h, proto, ld)
        return self._plans[c]

    def iterate(self, source, T):
        """T iterations of u <- D^-1 source + P u from u = 0.  source: (n, c) float64 (full, on every rank).
        Returns the (n, c) float64 result (full, on every rank)."""
        torch, gdev = self._torch, self._gdev
        from . import _lib
        source = np.asarray(source, dtype=np.float64)
        c = source.shape[1]
        h, proto, ld = self._plans.get(c) or self._setup(c)
        dev = proto.slab.device
        Db = torch.zeros((proto.rows_pad, ld), dtype=torch.float32, device=dev)
        loc = ((1.0 / self.deg[proto.r0:proto.r1])[:, None] * source[proto.r0:proto.r1]).astype(np.float32)     # D * source, ssl.py:636
        Db[: loc.shape[0], :c] = torch.from_numpy(loc).to(dev)
        proto.u[0].zero_(); proto.u[1].zero_()

        def local_step(u_full, out_slab):
            _lib.call("glb_poisson_step", h, gdev.ptr(Db), gdev.ptr(u_full), gdev.ptr(out_slab), gdev.cur_stream())

        self.launches = T
        out = proto.run(local_step, T)
        pos = torch.from_numpy(proto.padded_index()).to(dev)
        return out[pos, :c].double().cpu().numpy()

    def timed_iterations(self, c, T):
        """Device time (ms, CUDA events on the current stream) of T iterations on zeros; for bench.py."""
        torch, gdev = self._torch, self._gdev
        from . import _lib
        h, proto, ld = self._plans.get(c) or self._setup(c)
        Db = torch.zeros((proto.rows_pad, ld), dtype=torch.float32, device=proto.slab.device)

        def local_step(u_full, out_slab):
            _lib.call("glb_poisson_step", h, gdev.ptr(Db), gdev.ptr(u_full), gdev.ptr(out_slab), gdev.cur_stream())

        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        proto.run(local_step, T)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
