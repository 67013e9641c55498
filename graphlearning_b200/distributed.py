"""Row-partitioned Poisson iterate across the GPUs of one node (BASELINE config 5).

One process per GPU (torchrun).  The nodes are relabelled with a locality ordering (reverse Cuthill-McKee, integers
only), rank g owns a contiguous block of rows of P = D^-1 W^T in that numbering and the same rows of Db and u.  One
iteration of the reference loop `u = Db + P*u` (graphlearning/ssl.py:668) is, on every rank,

    rows of u_{t+1} that a neighbour needs   <-  Db + P u_t, written locally AND into the neighbours' label
                                                 matrices over NVLink by the kernel that computes them
    all other rows of u_{t+1}                <-  Db + P u_t, while those rows are in flight

(`PartitionedPoisson`, csrc/slab.cu, glb_slab_*): only the halo - the rows a rank's columns actually point to -
crosses the links, there is no collective on the data path and no pack/unpack pass.  The north star's variant
"one NCCL all-gather of the n x c label matrix per iteration" is kept as `AllGatherPoisson` (step kernel +
dist.all_gather_into_tensor): it is the baseline the halo exchange is measured against.  Both give bitwise the result
of a single-GPU run (same per-row arithmetic in the same order, no cross-rank reduction).

`HaloPartition` is the host-side structure (who owns what, what each rank receives and sends, local index spaces);
it is pure numpy and is what the world-size-2/3 gloo tests on CPU exercise.
"""
from __future__ import annotations

import ctypes

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


def locality_order(W):
    """perm[new] = old: the library's reverse Cuthill-McKee on the pattern of W (host, integers only)."""
    from . import _lib
    W = sparse.csr_matrix(W)
    rp = np.ascontiguousarray(W.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(W.indices, dtype=np.int32)
    perm = np.empty(W.shape[0], dtype=np.int32)
    _lib.call("glb_locality_order_host", ctypes.c_void_p(rp.ctypes.data), ctypes.c_void_p(ci.ctypes.data), W.shape[0],
              ctypes.c_void_p(perm.ctypes.data))
    return perm


def poisson_slab(W, perm, r0, r1):
    """Rows [r0, r1) of P = D^-1 W^T (ssl.py:615-616, 634-635) in the numbering perm[new] = old, built from W without
    forming the whole of P: (rowptr int32, global new column ids int32 ascending, fp32 values, fp64 degrees of the rows)."""
    W = sparse.csr_matrix(W)
    n = W.shape[0]
    perm = np.arange(n, dtype=np.int64) if perm is None else np.asarray(perm, dtype=np.int64)
    iperm = np.empty(n, dtype=np.int64)
    iperm[perm] = np.arange(n)
    deg = np.asarray(W.sum(axis=1)).ravel() - W.diagonal()            # degrees of W - diag(W)
    own_old = perm[r0:r1]
    S = sparse.csr_matrix(W[:, own_old].T)                            # rows = columns of W = rows of W^T, old column ids
    S = sparse.csr_matrix((S.data, iperm[S.indices], S.indptr), shape=(r1 - r0, n))
    S.sort_indices()
    rows = np.repeat(np.arange(r1 - r0), np.diff(S.indptr))
    keep = S.indices != rows + r0                                     # W - diag(W)
    with np.errstate(divide="ignore"):
        dinv = 1.0 / deg[own_old]                                     # degree_matrix(p=-1); isolated nodes give inf as in the reference
    val = (dinv[rows] * S.data)[keep]
    col = S.indices[keep]
    rp = np.concatenate(([0], np.cumsum(np.bincount(rows[keep], minlength=r1 - r0)))).astype(np.int32)
    return rp, col.astype(np.int32), val.astype(np.float32), deg[own_old]


class HaloPartition:
    """What one rank needs to know about the partition.  `allgather(obj) -> [obj of rank 0, ..., obj of rank world-1]`
    is the only communication (torch.distributed.all_gather_object, or a list for in-process tests).

    Local index space: own rows 0..m-1, then the halo rows (sorted by global id).  Attributes:
      rp, col, val        the slab in local indices (order of the entries inside a row untouched)
      halo                global ids of the halo rows;  recv_from[h] = how many of them rank h owns
      boundary            uint8 per own row: a peer needs it, or it reads a halo row
      send_ptr/peer/dst   CSR over the own rows: the puts that deliver a row (peer rank, row index in the peer's space)
      rows_total[h]       m_h + H_h + 1 of every rank (label-buffer rows);  neighbours: bit mask of exchange partners
    """

    def __init__(self, rp, col_global, val, bounds, rank, allgather):
        bounds = np.asarray(bounds, dtype=np.int64)
        world = len(bounds) - 1
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        m = r1 - r0
        col_global = np.asarray(col_global, dtype=np.int64)
        outside = (col_global < r0) | (col_global >= r1)
        halo = np.unique(col_global[outside])
        col = np.where(outside, m + np.searchsorted(halo, col_global), col_global - r0).astype(np.int32)
        owner = np.searchsorted(bounds, halo, side="right") - 1
        need = [halo[owner == h] for h in range(world)]                # what this rank gathers from rank h
        everyone = allgather({"need": need, "m": m, "H": len(halo)})
        self.rank, self.world, self.bounds, self.m = rank, world, bounds, m
        self.rp = np.ascontiguousarray(rp, dtype=np.int32)
        self.col, self.val, self.halo = col, np.ascontiguousarray(val, dtype=np.float32), halo
        self.rows_total = np.array([e["m"] + e["H"] + 1 for e in everyone], dtype=np.int64)
        self.recv_from = np.array([len(x) for x in need], dtype=np.int64)
        # puts: rank h gathers the rows everyone[h]["need"][rank] from here; they sit in h's halo behind the rows h
        # gathers from the ranks below this one (the halo is sorted by global id and ranks own ascending ranges)
        rows, peers, dsts = [], [], []
        for h in range(world):
            if h == rank:
                continue
            want = np.asarray(everyone[h]["need"][rank], dtype=np.int64)
            if len(want) == 0:
                continue
            first = everyone[h]["m"] + sum(len(everyone[h]["need"][g]) for g in range(rank))
            rows.append(want - r0)
            peers.append(np.full(len(want), h, dtype=np.int32))
            dsts.append((first + np.arange(len(want))).astype(np.int32))
        if rows:
            rows, peers, dsts = np.concatenate(rows), np.concatenate(peers), np.concatenate(dsts)
            order = np.argsort(rows, kind="stable")
            rows, peers, dsts = rows[order], peers[order], dsts[order]
        else:
            rows, peers, dsts = np.zeros(0, np.int64), np.zeros(0, np.int32), np.zeros(0, np.int32)
        self.send_ptr = np.concatenate(([0], np.cumsum(np.bincount(rows, minlength=m)))).astype(np.int64)
        self.send_peer, self.send_dst = peers, dsts
        reads_halo = np.zeros(m, dtype=bool)
        reads_halo[np.repeat(np.arange(m), np.diff(self.rp))[outside]] = True
        self.boundary = (reads_halo | (np.diff(self.send_ptr) > 0)).astype(np.uint8)
        mask = 0
        for h in range(world):
            if h != rank and (self.recv_from[h] > 0 or len(everyone[h]["need"][rank]) > 0):
                mask |= 1 << h
        self.neighbours = mask

    # ---- numpy emulation of one exchange-and-step (the CPU tests run the protocol with it) -----------------------
    def step_numpy(self, Db, u_local):
        """u_local: (m + H + 1, c) with valid own and halo rows -> the own rows of Db + P u, float64."""
        P = sparse.csr_matrix((self.val.astype(np.float64), self.col, self.rp), shape=(self.m, len(u_local)))
        return Db + P @ u_local

    def puts(self, new_rows):
        """{peer: (destination rows in the peer's space, values)} for freshly computed own rows."""
        out = {}
        src = np.repeat(np.arange(self.m), np.diff(self.send_ptr))
        for h in np.unique(self.send_peer):
            sel = self.send_peer == h
            out[int(h)] = (self.send_dst[sel], new_rows[src[sel]])
        return out


class PartitionedPoisson:
    """Device side of the halo-exchange iterate: one glb_slab per rank, label matrices in peer-mapped regions.
    Needs torch.distributed initialised (nccl, or nothing for world = 1) and one GPU per rank."""

    def __init__(self, W, rank=None, world=None, reorder=True, c=10):
        import torch
        from . import _lib, device as gdev
        self._torch, self._lib, self._gdev = torch, _lib, gdev
        if world is None or rank is None:
            import torch.distributed as dist
            rank, world = dist.get_rank(), dist.get_world_size()
        self.rank, self.world, self.c = rank, world, c
        W = sparse.csr_matrix(W)
        self.n = W.shape[0]
        self.perm = locality_order(W) if reorder else None
        order = np.arange(self.n) if self.perm is None else self.perm
        lens = np.bincount(W.indices, minlength=self.n)[order]         # row lengths of W^T in the new numbering
        self.bounds = partition_rows(np.concatenate(([0], np.cumsum(lens))), world)
        r0, r1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        rp, colg, val, deg = poisson_slab(W, self.perm, r0, r1)
        self.part = HaloPartition(rp, colg, val, self.bounds, rank, self._allgather)
        self.own_old = order[r0:r1]
        self.nnz_local = len(val)
        p = self.part
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.deg = torch.from_numpy(np.ascontiguousarray(deg)).to(dev)
        vp = ctypes.c_void_p
        self._slab = vp()
        send_peer = np.ascontiguousarray(p.send_peer, dtype=np.int32)
        send_dst = np.ascontiguousarray(p.send_dst, dtype=np.int32)
        _lib.call("glb_slab_create", ctypes.byref(self._slab), vp(p.rp.ctypes.data), vp(p.col.ctypes.data), vp(p.val.ctypes.data),
                  max(p.m, 1) if p.m else 1, len(p.halo), c, vp(p.boundary.ctypes.data) if p.m else None,
                  vp(p.send_ptr.ctypes.data) if p.m else None, vp(send_peer.ctypes.data), vp(send_dst.ctypes.data), gdev.cur_stream())
        self.ld = int(_lib.load().glb_slab_ld(self._slab))
        self.tile_slices = int(_lib.load().glb_slab_tile_slices(self._slab))
        rows_total = int(_lib.load().glb_slab_rows(self._slab))
        assert rows_total == int(p.rows_total[rank]), (rows_total, p.rows_total)
        # peer-mapped regions
        nbytes = int(_lib.load().glb_slab_region_bytes(self._slab))
        self._region = vp()
        handle = (ctypes.c_ubyte * 64)()
        _lib.call("glb_ipc_alloc", nbytes, ctypes.byref(self._region), handle)
        handles = self._allgather(bytes(handle))
        self._peer_ptrs = {}
        base = (vp * world)()
        for h in range(world):
            if h == rank:
                base[h] = self._region
            elif (p.neighbours >> h) & 1:
                q = vp()
                hb = (ctypes.c_ubyte * 64).from_buffer_copy(handles[h])
                _lib.call("glb_ipc_open", hb, ctypes.byref(q))
                self._peer_ptrs[h] = q
                base[h] = q
        rows_arr = np.ascontiguousarray(p.rows_total, dtype=np.int64)
        _lib.call("glb_slab_attach", self._slab, rank, world, base, vp(rows_arr.ctypes.data), ctypes.c_uint32(p.neighbours))
        self.region_bytes = nbytes
        self.launches = 0

    def _allgather(self, obj):
        if self.world == 1:
            return [obj]
        import torch.distributed as dist
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def _barrier(self):
        self._torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def close(self):
        lib = self._lib.load()
        self._barrier()
        for q in self._peer_ptrs.values():
            lib.glb_ipc_close(q)
        self._peer_ptrs = {}
        self._barrier()                                        # nobody still maps a region that is about to be freed
        if self._slab:
            lib.glb_slab_destroy(self._slab)
            self._slab = None
        if self._region:
            lib.glb_ipc_free(self._region)
            self._region = None

    def halo_bytes_per_iteration(self):
        """bytes this rank puts into its peers per iteration (row stride ld floats)."""
        return int(len(self.part.send_peer)) * self.ld * 4

    def _run(self, Db, T):
        gdev, lib = self._gdev, self._lib
        which, nl = ctypes.c_int(0), ctypes.c_int(0)
        lib.call("glb_slab_reset", self._slab, gdev.cur_stream())
        self._barrier()                                        # every rank's buffers are zero before anyone puts into them
        lib.call("glb_slab_iterate", self._slab, gdev.ptr(Db), int(T), ctypes.byref(which), ctypes.byref(nl), gdev.cur_stream())
        self.launches = nl.value
        return which.value

    def iterate(self, source, T):
        """T iterations of u <- D^-1 source + P u from u = 0.  source: (n, c) float64 in the caller's numbering (full, on
        every rank).  Returns the (n, c) float64 result (full, on every rank)."""
        torch, gdev, lib = self._torch, self._gdev, self._lib
        source = np.asarray(source, dtype=np.float64)
        assert source.shape == (self.n, self.c)
        m = self.part.m
        src = torch.from_numpy(np.ascontiguousarray(source[self.own_old])).to(self.dev)
        Db = torch.zeros((max(m, 1), self.ld), dtype=torch.float32, device=self.dev)
        if m:
            lib.call("glb_slab_pack", self._slab, gdev.ptr(src), gdev.ptr(self.deg), gdev.ptr(Db), gdev.cur_stream())
        which = self._run(Db, T)
        out = torch.empty((max(m, 1), self.c), dtype=torch.float64, device=self.dev)
        if m:
            lib.call("glb_slab_unpack", self._slab, which, gdev.ptr(out), gdev.cur_stream())
        lib.call("glb_slab_check", self._slab, gdev.cur_stream())
        self._barrier()
        mine = out[:m].cpu().numpy()
        parts = self._allgather((self.own_old, mine))
        u = np.empty((self.n, self.c), dtype=np.float64)
        for rows, vals in parts:
            u[rows] = vals
        return u

    def timed_iterations(self, T):
        """Device time (ms, CUDA events on the current stream) of T iterations with a zero source; for bench.py."""
        torch, gdev, lib = self._torch, self._gdev, self._lib
        Db = torch.zeros((max(self.part.m, 1), self.ld), dtype=torch.float32, device=self.dev)
        which, nl = ctypes.c_int(0), ctypes.c_int(0)
        lib.call("glb_slab_reset", self._slab, gdev.cur_stream())
        self._barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.call("glb_slab_iterate", self._slab, gdev.ptr(Db), int(T), ctypes.byref(which), ctypes.byref(nl), gdev.cur_stream())
        e1.record()
        torch.cuda.synchronize()
        lib.call("glb_slab_check", self._slab, gdev.cur_stream())
        self.launches = nl.value
        return e0.elapsed_time(e1)


# -------------------------------------------------------------------------------------------------------------------------
# baseline: one all-gather of the whole label matrix per iteration
# -------------------------------------------------------------------------------------------------------------------------
class PartitionedIterate:
    """The all-gather protocol, independent of where the local product runs.

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


class AllGatherPoisson:
    """Baseline of SURVEY 8(e): row slab of P on this rank's GPU, step kernel + one NCCL all-gather of the n x c label
    matrix per iteration.  Needs torch.distributed initialised with the nccl backend and one GPU per rank."""

    def __init__(self, W, rank=None, world=None, reorder=False):
        import torch
        import torch.distributed as dist
        from . import device as gdev
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        W = sparse.csr_matrix(W)
        self.n = W.shape[0]
        self.perm = locality_order(W) if reorder else None
        order = np.arange(self.n) if self.perm is None else self.perm
        lens = np.bincount(W.indices, minlength=self.n)[order]
        self.bounds = partition_rows(np.concatenate(([0], np.cumsum(lens))), self.world)
        r0, r1 = int(self.bounds[self.rank]), int(self.bounds[self.rank + 1])
        rp, col, val, deg = poisson_slab(W, self.perm, r0, r1)         # this rank's rows only
        if r1 == r0:
            rp = np.zeros(2, dtype=np.int32)
        self.deg_local, self.order = deg, order
        self._slab_host = (rp, col, val)
        self.nnz_local = len(col)
        self._torch, self._dist, self._gdev = torch, dist, gdev
        self._plans = {}

    def _setup(self, c):
        torch, dist, gdev = self._torch, self._dist, self._gdev
        from . import _lib
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
        own_old = self.order[proto.r0:proto.r1]
        with np.errstate(divide="ignore"):
            loc = ((1.0 / self.deg_local)[:, None] * source[own_old]).astype(np.float32)     # D * source, ssl.py:636
        Db[: loc.shape[0], :c] = torch.from_numpy(loc).to(dev)
        proto.u[0].zero_(); proto.u[1].zero_()

        def local_step(u_full, out_slab):
            _lib.call("glb_poisson_step", h, gdev.ptr(Db), gdev.ptr(u_full), gdev.ptr(out_slab), gdev.cur_stream())

        self.launches = T
        out = proto.run(local_step, T)
        pos = torch.from_numpy(proto.padded_index()).to(dev)
        u_new = out[pos, :c].double().cpu().numpy()                   # rows in the relabelled numbering
        u = np.empty_like(u_new)
        u[self.order] = u_new
        return u

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
