"""Import the UNMODIFIED reference (jwcalder/GraphLearning) from /root/reference under the
alias ``gl_ref`` so that golden vectors can be generated in the build container.

TEST INFRASTRUCTURE ONLY - used by oracle/make_golden.py (run by hand in the build
container, never on the GPU box: /root/reference does not exist there).

The reference imports matplotlib at module top (graphlearning/ssl.py:119-120,
graph.py:11, utils.py:12); matplotlib is not installed here, so empty stand-in modules
are seeded first.  The package object is created by hand because
graphlearning/__init__.py:1-8 uses absolute imports of its own name.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("GL_REFERENCE_ROOT", "/root/reference")


def load_reference():
    if "gl_ref" in sys.modules:
        return sys.modules["gl_ref"]
    pkgdir = os.path.join(REF_ROOT, "graphlearning")
    if not os.path.isdir(pkgdir):
        raise RuntimeError("reference checkout not found at " + REF_ROOT)
    for m in ("matplotlib", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = types.ModuleType(m)
    sys.modules["matplotlib"].rcParams = {}
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    pkg = types.ModuleType("gl_ref")
    pkg.__path__ = [pkgdir]
    sys.modules["gl_ref"] = pkg
    for sub in ("utils", "graph", "weightmatrix", "ssl"):
        setattr(pkg, sub, importlib.import_module("gl_ref." + sub))
    return pkg
