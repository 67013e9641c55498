"""Build libglb200.so (the sm_100a CUDA kernels + C-ABI) in-tree with nvcc.

    python -m graphlearning_b200.build [--force] [--verbose] [--exp]

The library lands in graphlearning_b200/lib/libglb200.so (git-ignored, but it travels to the GPU box
with the gpurun snapshot).  nvcc cross-compiles without a GPU.

--exp builds lib/libglb200_exp.so with -DGLB_EXPERIMENT: the same sources with the experiment switches
(GLB_POISSON_* environment variables, read once per plan) compiled in.  Only tools/ load it (GLB200_LIB=...);
the product library has no environment lookups.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libglb200.so")
OBJDIR = os.path.join(HERE, "build")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


LIB_EXP = os.path.join(LIBDIR, "libglb200_exp.so")


def needs_build(lib=LIB):
    return not os.path.exists(lib) or os.path.getmtime(lib) < _deps_mtime()


def build(force=False, verbose=False, exp=False):
    lib = LIB_EXP if exp else LIB
    objdir = OBJDIR + ("_exp" if exp else "")
    # GLB_NVCC_EXTRA: more -D switches for the experiment build only (compile-time A/B runs of tools/)
    extra = ["-DGLB_EXPERIMENT"] + os.environ.get("GLB_NVCC_EXTRA", "").split() if exp else []
    if not force and not needs_build(lib):
        return lib
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(objdir, exist_ok=True)
    dep_m = _deps_mtime()

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= dep_m:
            return obj, ""
        cmd = [NVCC] + ARCH_FLAGS + CFLAGS + extra + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj, r.stderr

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, sources()))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        for _, log in results:
            f.write(log)
    cmd = [NVCC] + ARCH_FLAGS + ["-shared", "-o", lib] + objs + ["-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, exp="--exp" in sys.argv))
