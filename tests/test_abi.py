"""The C-ABI library loads and exports every symbol include/glb200.h declares (no compute, no GPU)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from graphlearning_b200 import build, _lib
    build.build()                       # nvcc cross-compiles without a GPU; no-op when up to date
    return _lib.load()


def declared_symbols():
    names = []
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            names += re.findall(r"GLB_API\s+[\w\s\*]+?\b(glb_\w+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_symbols():
    assert len(declared_symbols()) >= 15


def test_every_declared_symbol_is_exported(lib):
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_signatures_cover_header(lib):
    from graphlearning_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_version_and_padding(lib):
    assert lib.glb_version() >= 100
    assert [lib.glb_padded_ld(c) for c in (1, 2, 4, 5, 10, 16, 17, 50, 100, 128, 129, 300)] == \
        [4, 4, 4, 8, 16, 16, 32, 64, 128, 128, 256, 384]
    assert lib.glb_padded_ld(0) < 0


def test_argument_errors_are_reported_not_crashes(lib):
    rc = lib.glb_poisson_step(None, None, None, None, None)
    assert rc == -1
    assert b"null pointer" in lib.glb_last_error()
    h = ctypes.c_void_p()
    assert lib.glb_poisson_plan_create(ctypes.byref(h), None, None, None, 10, 10, 10, -1, None) == -1
    assert lib.glb_poisson_plan_kind(None) < 0 and lib.glb_poisson_plan_ld(None) < 0
    assert lib.glb_csr_transpose_work_bytes(10, -5) < 0


def test_no_silent_cpu_fallback(lib):
    """Without a device the host entry point must fail loudly (GLB_E_NOGPU), never compute on the CPU."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    import graphlearning_b200 as gl
    from graphlearning_b200._lib import GlbError
    from scipy import sparse
    W = sparse.random(50, 50, 0.2, format="csr", random_state=0)
    W = W + W.T
    with pytest.raises(GlbError, match="no CUDA device"):
        gl.ssl.poisson(W, solver="gradient_descent").fit(np.array([0, 1]), np.array([0, 1]))
