"""Build the C-ABI kernel library in-tree with nvcc for sm_100a.

    python -m diffmvs_b200.build [--force]

Produces `diffmvs_b200/lib/libdiffmvs_b200.so` (git-ignored; it travels to the GPU box with the
gpurun snapshot).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libdiffmvs_b200.so")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


LEGACY = os.environ.get("DMVS_BUILD_LEGACY", "0") not in ("", "0")


def sources():
    """Kernel sources of the shipped library.  `csrc/legacy/` (the round-1 mma.sync and tap-offset tcgen05 convolution
    back ends, off every shipped configuration) is compiled only with DMVS_BUILD_LEGACY=1."""
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if LEGACY:
        srcs += sorted(glob.glob(os.path.join(CSRC, "legacy", "*.cu")))
    return srcs


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(PKG, "build")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    common = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
              "-I", CSRC] + ARCH_FLAGS + (["-DDMVS_LEGACY_BACKENDS=1"] if LEGACY else [])
    if verbose:
        common += ["-Xptxas", "-v"]
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        procs.append((src, subprocess.Popen([nvcc, "-c", src, "-o", obj] + common, stdout=subprocess.PIPE,
                                            stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {os.path.basename(src)} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.check_call([nvcc, "-shared", "-o", LIB_PATH] + objs + ARCH_FLAGS)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
