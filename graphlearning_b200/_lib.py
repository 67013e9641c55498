"""ctypes binding of libglb200.so (C-ABI declared in include/glb200.h).

There is NO CPU fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# GLB200_LIB: another build of the same library (tools/ use lib/libglb200_exp.so, the -DGLB_EXPERIMENT build)
LIB_PATH = os.environ.get("GLB200_LIB") or os.path.join(HERE, "lib", "libglb200.so")

_lib = None

# name -> (restype, argtypes).  Mirrors include/glb200.h one to one (tests/test_abi.py checks that).
SIGNATURES = {
    "glb_version": (c_int, []),
    "glb_release_workspace": (c_int, []),
    "glb_last_error": (c_char_p, []),
    "glb_device_info": (c_int, [POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int64)]),
    "glb_padded_ld": (c_int, [c_int]),
    "glb_csr_degree": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p]),
    "glb_csr_transpose_work_bytes": (c_int64, [c_int64, c_int64]),
    "glb_csr_transpose": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int64, c_void_p]),
    "glb_poisson_scale": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "glb_locality_order_host": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "glb_csr_permute": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p]),
    "glb_poisson_plan_create": (c_int, [POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                        c_void_p]),
    "glb_poisson_plan_destroy": (c_int, [c_void_p]),
    "glb_poisson_plan_kind": (c_int, [c_void_p]),
    "glb_poisson_plan_ld": (c_int, [c_void_p]),
    "glb_poisson_plan_rows": (c_int64, [c_void_p]),
    "glb_poisson_plan_check": (c_int, [c_void_p, c_void_p]),
    "glb_poisson_plan_fill": (c_double, [c_void_p]),
    "glb_poisson_plan_gate": (c_int, [c_void_p]),
    "glb_laplacian_csr_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "glb_laplace_fit_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_int64, c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "glb_dataflow_slabs_check_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p]),
    "glb_poisson_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "glb_poisson_unpack": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "glb_poisson_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "glb_poisson_iterate": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, POINTER(c_int), POINTER(c_int),
                                    c_void_p]),
    "glb_poisson_mixing_T": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int,
                                     POINTER(c_int), POINTER(c_int), c_void_p]),
    "glb_poisson_graph_create": (c_int, [POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int]),
    "glb_poisson_graph_destroy": (c_int, [c_void_p]),
    "glb_poisson_graph_fit": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int64, c_int, c_int, c_void_p,
                                      POINTER(c_int), POINTER(c_int)]),
    "glb_laplace_graph_create": (c_int, [POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p,
                                         c_void_p, c_void_p]),
    "glb_laplace_graph_fit": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int, c_double, c_void_p, POINTER(c_int64),
                                      POINTER(c_double), POINTER(c_int), c_void_p]),
    "glb_laplace_graph_destroy": (c_int, [c_void_p]),
    "glb_poisson_graph_fit_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int64, c_int, c_int,
                                           c_void_p, POINTER(c_int), POINTER(c_int)]),
    "glb_host_alloc": (c_int, [c_int64, POINTER(c_void_p)]),
    "glb_host_free": (c_int, [c_void_p]),
    "glb_knn_search": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, POINTER(c_int), POINTER(c_int), c_void_p]),
    "glb_knn_search_host": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, POINTER(c_int), POINTER(c_int)]),
    "glb_cg_work_bytes": (c_int64, [c_int64, c_int]),
    "glb_cg_solve": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_double, c_int64,
                             c_void_p, c_void_p, c_int64, POINTER(c_int64), POINTER(c_double), POINTER(c_int), c_void_p]),
    "glb_cg_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int, c_double, c_int64,
                            c_void_p, POINTER(c_int64), POINTER(c_double), POINTER(c_int)]),
    "glb_lp_iterate_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int,
                                    c_double, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "glb_lip_iterate_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_double, c_int,
                                     c_double, c_double, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "glb_spmm_f64": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int, c_int, c_double,
                             c_void_p, c_int, c_double, c_void_p, c_void_p, c_int, c_double, c_void_p]),
    "glb_gram_work_bytes": (c_int64, [c_int, c_int]),
    "glb_gram_f64": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p, c_int64, c_void_p]),
    "glb_right_mul_f64": (c_int, [c_void_p, c_int, c_int64, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p]),
    "glb_knn_weights_csr_host": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64,
                                         POINTER(c_int64), POINTER(c_int)]),
    "glb_lip_iterate_multi_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_double, c_int,
                                           c_double, c_double, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "glb_slab_create": (c_int, [POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "glb_slab_destroy": (c_int, [c_void_p]),
    "glb_slab_rows": (c_int64, [c_void_p]),
    "glb_slab_ld": (c_int, [c_void_p]),
    "glb_slab_check_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p, c_int, c_void_p]),
    "glb_slab_fill": (c_double, [c_void_p]),
    "glb_slab_tile_slices": (c_int, [c_void_p]),
    "glb_slab_region_bytes": (c_int64, [c_void_p]),
    "glb_slab_attach": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, ctypes.c_uint32]),
    "glb_slab_buffer": (c_int, [c_void_p, c_int, POINTER(c_void_p)]),
    "glb_slab_pack": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "glb_slab_unpack": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "glb_slab_reset": (c_int, [c_void_p, c_void_p]),
    "glb_slab_iterate": (c_int, [c_void_p, c_void_p, c_int, POINTER(c_int), POINTER(c_int), c_void_p]),
    "glb_slab_check": (c_int, [c_void_p, c_void_p]),
    "glb_ipc_alloc": (c_int, [c_int64, POINTER(c_void_p), c_void_p]),
    "glb_ipc_open": (c_int, [c_void_p, POINTER(c_void_p)]),
    "glb_ipc_close": (c_int, [c_void_p]),
    "glb_ipc_free": (c_int, [c_void_p]),
    "glb_volume_projection": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p]),
    "glb_onehot_f64": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "glb_max_abs_diff_f64": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, POINTER(c_double), c_void_p]),
    "glb_centered_step_f64": (c_int, [c_void_p, c_int, c_void_p, c_double, c_void_p, c_int, c_void_p, c_int64, c_int,
                                       POINTER(c_double), c_void_p]),
    "glb_min_nonneg_f64": (c_int, [c_void_p, c_int64, c_int, c_int, POINTER(c_double), c_void_p]),
    "glb_argmax_rows_f64": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p]),
    "glb_poisson_gd_host": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_int64,
                                    c_int, c_int, c_void_p, POINTER(c_int), POINTER(c_int)]),
}


class GlbError(RuntimeError):
    pass


def load():
    """Load libglb200.so; raises RuntimeError (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libglb200.so not found at %s - build it with `python -m graphlearning_b200.build` "
                           "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().glb_last_error()
        raise GlbError("%s failed (code %d): %s" % (what or "libglb200 call", rc, msg.decode() if msg else ""))


def call(name, *args):
    check(getattr(load(), name)(*args), name)


def device_info():
    sm, maj, mnr, mem = c_int(), c_int(), c_int(), c_int64()
    call("glb_device_info", ctypes.byref(sm), ctypes.byref(maj), ctypes.byref(mnr), ctypes.byref(mem))
    return dict(sm_count=sm.value, cc=(maj.value, mnr.value), hbm_bytes=mem.value)


def padded_ld(c):
    ld = load().glb_padded_ld(int(c))
    if ld <= 0:
        raise GlbError("glb_padded_ld(%r) rejected" % (c,))
    return ld
