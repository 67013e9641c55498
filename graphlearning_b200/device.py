"""Device-resident building blocks: CSR graphs and the Poisson operator held in HBM.

torch is used only as the owner of device memory and streams (tensor.data_ptr(), current stream); every
kernel is launched through the C-ABI of libglb200.so.
"""
from __future__ import annotations

import ctypes

import numpy as np
from scipy import sparse

from . import _lib


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("graphlearning_b200 needs a CUDA device (B200); there is no CPU fallback")
    return torch


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def cur_stream():
    return ctypes.c_void_p(_torch().cuda.current_stream().cuda_stream)


class DeviceCSR:
    """A CSR matrix in HBM: int32 rowptr/col, values fp64 (setup) and/or fp32 (iterate)."""

    def __init__(self, rowptr, col, val, n):
        self.rowptr, self.col, self.val, self.n = rowptr, col, val, int(n)
        self.nnz = int(col.numel())

    @classmethod
    def from_scipy(cls, W, device="cuda"):
        torch = _torch()
        W = sparse.csr_matrix(W)
        if W.shape[0] != W.shape[1]:
            raise ValueError("weight matrix must be square")
        rp = torch.from_numpy(np.ascontiguousarray(W.indptr, dtype=np.int32)).to(device)
        col = torch.from_numpy(np.ascontiguousarray(W.indices, dtype=np.int32)).to(device)
        val = torch.from_numpy(np.ascontiguousarray(W.data, dtype=np.float64)).to(device)
        return cls(rp, col, val, W.shape[0])

    def to_scipy(self):
        return sparse.csr_matrix((self.val.cpu().numpy(), self.col.cpu().numpy(), self.rowptr.cpu().numpy()),
                                 shape=(self.n, self.n))

    def degree(self, skip_diagonal=False):
        """graph.degree_vector (reference graphlearning/graph.py:108-122) on the device."""
        torch = _torch()
        deg = torch.empty(self.n, dtype=torch.float64, device=self.val.device)
        _lib.call("glb_csr_degree", ptr(self.rowptr), ptr(self.col), ptr(self.val), self.n, int(skip_diagonal),
                  ptr(deg), cur_stream())
        return deg

    def transpose(self):
        torch = _torch()
        dev = self.val.device
        wb = _lib.load().glb_csr_transpose_work_bytes(self.n, self.nnz)
        work = torch.empty(int(wb), dtype=torch.uint8, device=dev)
        t_rp = torch.empty(self.n + 1, dtype=torch.int32, device=dev)
        t_col = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)[: self.nnz]
        t_val = torch.empty(max(self.nnz, 1), dtype=torch.float64, device=dev)[: self.nnz]
        _lib.call("glb_csr_transpose", ptr(self.rowptr), ptr(self.col), ptr(self.val), self.n, self.nnz, ptr(t_rp),
                  ptr(t_col), ptr(t_val), ptr(work), int(wb), cur_stream())
        return DeviceCSR(t_rp, t_col, t_val, self.n)


class PoissonOperator:
    """P = D^-1 W^T (fp32 CSR) and RW = W^T D^-1 (fp64 values on the same pattern), device resident.
    Setup lines of ssl.poisson._fit, reference graphlearning/ssl.py:615-617, 634-644."""

    KINDS = {"auto": -1, "step": 0, "barrier": 1, "dataflow": 2}

    def __init__(self, W, reorder=False, kind="auto"):
        """W: scipy CSR (host) or DeviceCSR.  reorder=True relabels the nodes with a locality ordering
        (reverse Cuthill-McKee from the host pattern) for the iterate; inputs/outputs keep the caller's
        numbering.  Only possible when W arrives as a host matrix.  kind: which iterate kernel the plans
        of this operator use ("auto" picks dataflow -> barrier -> step, see include/glb200.h)."""
        torch = _torch()
        host_W = None if isinstance(W, DeviceCSR) else sparse.csr_matrix(W)
        self.W = W if isinstance(W, DeviceCSR) else DeviceCSR.from_scipy(host_W)
        self.n = self.W.n
        self.kind_request = self.KINDS[kind] if isinstance(kind, str) else int(kind)
        self.deg = self.W.degree(skip_diagonal=True)
        Wt = self.W.transpose()
        self.rowptr, self.col, self.nnz = Wt.rowptr, Wt.col, Wt.nnz
        self.P_val = torch.empty(max(self.nnz, 1), dtype=torch.float32, device=self.deg.device)
        self.RW_val = torch.empty(max(self.nnz, 1), dtype=torch.float64, device=self.deg.device)
        _lib.call("glb_poisson_scale", ptr(Wt.rowptr), ptr(Wt.col), ptr(Wt.val), ptr(self.deg), self.n,
                  ptr(self.P_val), ptr(self.RW_val), cur_stream())
        # the arrays the iterate reads (relabelled when a locality ordering is in use)
        self.perm = None
        self.it_rowptr, self.it_col, self.it_val = self.rowptr, self.col, self.P_val
        if reorder and host_W is not None and self.nnz > 0:
            rp = np.ascontiguousarray(host_W.indptr, dtype=np.int32)
            ci = np.ascontiguousarray(host_W.indices, dtype=np.int32)
            perm = np.empty(self.n, dtype=np.int32)
            _lib.call("glb_locality_order_host", ctypes.c_void_p(rp.ctypes.data), ctypes.c_void_p(ci.ctypes.data),
                      self.n, ctypes.c_void_p(perm.ctypes.data))
            dev = self.deg.device
            self.perm = torch.from_numpy(perm).to(dev)
            self._iperm = torch.empty(self.n, dtype=torch.int32, device=dev)
            self.it_rowptr = torch.empty(self.n + 1, dtype=torch.int32, device=dev)
            self.it_col = torch.empty(self.nnz, dtype=torch.int32, device=dev)
            self.it_val = torch.empty(self.nnz, dtype=torch.float32, device=dev)
            _lib.call("glb_csr_permute", ptr(self.rowptr), ptr(self.col), ptr(self.P_val), self.n, self.nnz,
                      ptr(self.perm), ptr(self._iperm), ptr(self.it_rowptr), ptr(self.it_col), ptr(self.it_val),
                      cur_stream())
        self._plans = {}
        self._last_c = None

    def __del__(self):
        try:
            for p in self._plans.values():
                _lib.load().glb_poisson_plan_destroy(p)
        except Exception:
            pass

    def plan(self, c=None):
        """The glb_poisson_plan for label matrices with c columns (built on first use)."""
        c = int(self._last_c if c is None else c)
        if c not in self._plans:
            h = ctypes.c_void_p()
            _lib.call("glb_poisson_plan_create", ctypes.byref(h), ptr(self.it_rowptr), ptr(self.it_col), ptr(self.it_val),
                      self.n, self.nnz, c, self.kind_request, cur_stream())
            self._plans[c] = h
        self._last_c = c
        return self._plans[c]

    def kind(self, c=None):
        """'step' | 'barrier' | 'dataflow': the kernel the plan for width c runs."""
        k = _lib.load().glb_poisson_plan_kind(self.plan(c))
        return {v: n for n, v in self.KINDS.items()}[k]

    def ld(self, c=None):
        return int(_lib.load().glb_poisson_plan_ld(self.plan(c)))

    def rows(self, c=None):
        """Rows of a device label matrix: n, plus the library's scratch row for the dataflow kernel."""
        return int(_lib.load().glb_poisson_plan_rows(self.plan(c)))

    def fill(self, c=None):
        return float(_lib.load().glb_poisson_plan_fill(self.plan(c)))

    def gate(self, c=None):
        return int(_lib.load().glb_poisson_plan_gate(self.plan(c)))

    def is_persistent(self, c=None):
        return self.kind(c) != "step"

    def pack(self, X, scale_by_degree=False):
        """host/device (n,c) float64 -> device (n,ld) fp32 in the layout of the plan for width c."""
        torch = _torch()
        X = torch.as_tensor(X, dtype=torch.float64).to(self.deg.device).contiguous()
        n, c = X.shape
        plan = self.plan(c)
        out = torch.zeros((self.rows(c), self.ld(c)), dtype=torch.float32, device=X.device)
        _lib.call("glb_poisson_pack", plan, ptr(X), ptr(self.deg) if scale_by_degree else None, ptr(self.perm), ptr(out),
                  cur_stream())
        return out

    def unpack(self, U, c=None):
        torch = _torch()
        plan = self.plan(c)
        c = self._last_c
        out = torch.empty((self.n, c), dtype=torch.float64, device=U.device)
        _lib.call("glb_poisson_plan_check", plan, cur_stream())        # raises if a polling loop hit its watchdog
        _lib.call("glb_poisson_unpack", plan, ptr(U), ptr(self.perm), ptr(out), cur_stream())
        return out

    def source_to_Db(self, source):
        """Db = D^-1 source (ssl.py:636) in the device layout."""
        return self.pack(source, scale_by_degree=True)

    def step(self, Db, u_in, u_out, c=None):
        _lib.call("glb_poisson_step", self.plan(c), ptr(Db), ptr(u_in), ptr(u_out), cur_stream())

    def iterate(self, Db, T, u0=None, u1=None, c=None):
        """T iterations of u <- Db + P u from u0 (zeros by default).  Returns (u, launches)."""
        torch = _torch()
        if u0 is None:
            u0 = torch.zeros_like(Db)
        if u1 is None:
            u1 = torch.zeros_like(Db)
        which, launches = ctypes.c_int(0), ctypes.c_int(0)
        _lib.call("glb_poisson_iterate", self.plan(c), ptr(Db), ptr(u0), ptr(u1), int(T), ctypes.byref(which),
                  ctypes.byref(launches), cur_stream())
        return (u1 if which.value else u0), launches.value

    def mixing_T(self, train_ind, min_iter, max_iter):
        """Iteration count by the reference's stopping rule (ssl.py:639-644, 667, 669)."""
        torch = _torch()
        dev = self.deg.device
        v = torch.zeros(self.n, dtype=torch.float64, device=dev)
        v[torch.as_tensor(np.asarray(train_ind), dtype=torch.long, device=dev)] = 1
        v = v / v.sum()
        vinf = self.deg / self.deg.sum()
        tmp = torch.empty_like(v)
        T, launches = ctypes.c_int(0), ctypes.c_int(0)
        _lib.call("glb_poisson_mixing_T", ptr(self.rowptr), ptr(self.col), ptr(self.RW_val), ptr(vinf), ptr(v), ptr(tmp),
                  self.n, int(min_iter), int(max_iter), ctypes.byref(T), ctypes.byref(launches), cur_stream())
        return T.value


class _PinnedPool:
    """Result arrays in page-locked host memory (glb_host_alloc): the download of an n x c score matrix into one runs at
    PCIe speed, into a fresh pageable numpy array it pays a staging copy and a page fault per 4 KB (1 ms for 5.6 MB).
    Page-locking itself is expensive (3-16 ms for 5.6 MB, profiles/r2_pinned_alloc_probe.txt), so it never happens on the
    caller's thread: a request that finds no idle buffer of its size gets an ordinary numpy array and starts a background
    thread that pins two buffers of that size (the result a model still holds + the next one) for the fits to come.  A
    buffer goes back to the pool when the array (and every view of it) is garbage collected; at most `keep` idle buffers
    per size stay pinned and at most `cap_bytes` in total."""

    def __init__(self, keep=4, cap_bytes=1 << 30):
        import threading
        self.keep, self.cap_bytes = keep, cap_bytes
        self.idle, self.pending, self.total = {}, {}, 0
        self.lock = threading.Lock()

    def empty(self, shape, dtype=np.float64):
        import weakref
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        addr = None
        if nbytes:
            with self.lock:
                lst = self.idle.get(nbytes)
                if lst:
                    addr = lst.pop()
        if addr is None:
            if nbytes >= (1 << 16):
                try:
                    self._prefetch(nbytes, 2)
                except Exception:                                  # the pool is an optimisation: never fail a request over it
                    pass
            return np.empty(shape, dtype=dtype)
        flat = np.frombuffer((ctypes.c_char * nbytes).from_address(addr), dtype=dtype)
        weakref.finalize(flat, self._release, nbytes, addr)       # views keep `flat` alive through .base
        return flat.reshape(shape)

    def _prefetch(self, nbytes, count):
        import threading
        with self.lock:
            have = len(self.idle.get(nbytes, ())) + self.pending.get(nbytes, 0)
            count = min(count - have, (self.cap_bytes - self.total) // nbytes)
            if count <= 0:
                return
            self.pending[nbytes] = self.pending.get(nbytes, 0) + count
            self.total += count * nbytes
        dev = _torch().cuda.current_device()

        def work():
            for _ in range(count):
                p, ok = ctypes.c_void_p(), False
                try:
                    _torch().cuda.set_device(dev)                  # the thread's own current device
                    ok = _lib.load().glb_host_alloc(ctypes.c_int64(nbytes), ctypes.byref(p)) == 0 and bool(p.value)
                except Exception:
                    ok = False
                with self.lock:
                    self.pending[nbytes] -= 1
                    if ok:
                        self.idle.setdefault(nbytes, []).append(p.value)
                    else:
                        self.total -= nbytes
        threading.Thread(target=work, daemon=True).start()

    def wait(self, timeout=None):
        """Block until the background allocations are done (tests, benchmarks; at interpreter exit with a timeout, so that
        no thread is inside the driver when CUDA is torn down)."""
        import time
        t0 = time.monotonic()
        while True:
            with self.lock:
                if not any(self.pending.values()):
                    return
            if timeout is not None and time.monotonic() - t0 > timeout:
                return
            time.sleep(0.001)

    def _release(self, nbytes, addr):
        with self.lock:
            lst = self.idle.setdefault(nbytes, [])
            if len(lst) < self.keep:
                lst.append(addr)
                return
            self.total -= nbytes
        try:
            _lib.load().glb_host_free(ctypes.c_void_p(addr))
        except Exception:
            pass


pinned = _PinnedPool()
import atexit
atexit.register(pinned.wait, 5.0)


class PoissonGraphHandle:
    """Owner of a glb_poisson_graph (include/glb200.h): the device-resident P / RW / degree state of one weight
    matrix, reused by every fit on that graph (the reference rebuilds all of it inside each _fit call)."""

    def __init__(self, W, reorder=-1):
        W = sparse.csr_matrix(W)
        if W.shape[0] != W.shape[1]:
            raise ValueError("weight matrix must be square")
        self.n = W.shape[0]
        rp = np.ascontiguousarray(W.indptr, dtype=np.int32)
        col = np.ascontiguousarray(W.indices, dtype=np.int32)
        val = np.ascontiguousarray(W.data, dtype=np.float64)
        self._h = ctypes.c_void_p()
        _lib.call("glb_poisson_graph_create", ctypes.byref(self._h), ctypes.c_void_p(rp.ctypes.data),
                  ctypes.c_void_p(col.ctypes.data), ctypes.c_void_p(val.ctypes.data), self.n, len(col), int(reorder))
        self.h2d_bytes = rp.nbytes + col.nbytes + val.nbytes

    def fit(self, source, train_ind, min_iter, max_iter):
        """-> (u (n,c) float64, T, kernel launches)."""
        src = np.ascontiguousarray(source, dtype=np.float64)
        ti = np.ascontiguousarray(train_ind, dtype=np.int64)
        n, c = src.shape
        if n != self.n:
            raise ValueError("source has %d rows, graph has %d nodes" % (n, self.n))
        u = np.empty((n, c), dtype=np.float64)
        T, nl = ctypes.c_int(0), ctypes.c_int(0)
        _lib.call("glb_poisson_graph_fit", self._h, ctypes.c_void_p(src.ctypes.data), c, ctypes.c_void_p(ti.ctypes.data),
                  len(ti), int(min_iter), int(max_iter), ctypes.c_void_p(u.ctypes.data), ctypes.byref(T), ctypes.byref(nl))
        return u, T.value, nl.value

    def fit_rows(self, row_ind, rows, train_ind, min_iter, max_iter):
        """The same fit with the source term given by its nonzero rows (source = zeros((n, c)); source[row_ind] = rows,
        ssl.py:619-622): a few hundred bytes go up instead of n x c x 8, the scores come down into pinned memory.
        -> (u (n,c) float64, T, kernel launches)."""
        rows = np.ascontiguousarray(rows, dtype=np.float64)
        ri = np.ascontiguousarray(row_ind, dtype=np.int64).ravel()
        ti = np.ascontiguousarray(train_ind, dtype=np.int64).ravel()
        if rows.ndim != 2 or rows.shape[0] != len(ri):
            raise ValueError("rows must be (len(row_ind), c)")
        c = rows.shape[1]
        u = pinned.empty((self.n, c))
        T, nl = ctypes.c_int(0), ctypes.c_int(0)
        _lib.call("glb_poisson_graph_fit_rows", self._h, ctypes.c_void_p(ri.ctypes.data), ctypes.c_void_p(rows.ctypes.data),
                  len(ri), c, ctypes.c_void_p(ti.ctypes.data), len(ti), int(min_iter), int(max_iter),
                  ctypes.c_void_p(u.ctypes.data), ctypes.byref(T), ctypes.byref(nl))
        return u, T.value, nl.value

    def __del__(self):
        try:
            if self._h:
                _lib.load().glb_poisson_graph_destroy(self._h)
                self._h = None
        except Exception:
            pass


class LaplaceGraphHandle:
    """Owner of a glb_laplace_graph: the weight matrix and the scalings of L = Diag(diag) - Diag(left) W Diag(right) (+ tau)
    resident in HBM, shared by every Laplace-learning fit on the graph with that normalisation."""

    def __init__(self, rp, ci, val, n, left, right, diag, tau):
        self.n = int(n)
        self._h = ctypes.c_void_p()
        vp = lambda a: ctypes.c_void_p(a.ctypes.data) if a is not None else None
        _lib.call("glb_laplace_graph_create", ctypes.byref(self._h), vp(rp), vp(ci), vp(val), self.n, len(ci), vp(left), vp(right),
                  vp(diag), vp(tau))

    def fit(self, train_ind, F, tol):
        """-> (u (n,c) float64 in pinned memory, CG iterations, err, launches, (cg device ms, system nnz, unknowns))."""
        ti = np.ascontiguousarray(train_ind, dtype=np.int64)
        F = np.ascontiguousarray(F, dtype=np.float64)
        u = pinned.empty((self.n, F.shape[1]))
        it, err, nl = ctypes.c_int64(0), ctypes.c_double(0.0), ctypes.c_int(0)
        ms = np.zeros(3)
        vp = lambda a: ctypes.c_void_p(a.ctypes.data)
        _lib.call("glb_laplace_graph_fit", self._h, vp(ti), len(ti), vp(F), F.shape[1], float(tol), vp(u), ctypes.byref(it),
                  ctypes.byref(err), ctypes.byref(nl), vp(ms))
        return u, it.value, err.value, nl.value, ms

    def __del__(self):
        try:
            if self._h:
                _lib.load().glb_laplace_graph_destroy(self._h)
                self._h = None
        except Exception:
            pass
