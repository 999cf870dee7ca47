"""Builds libb200gs.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python 3d-gaussian-splatting-for-novel-view-synthesis_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU, so this runs in the build container; the resulting .so is
git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "b200gs", "libb200gs.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
          "--expt-relaxed-constexpr", "-I", INCLUDE]
# per-file extra flags; preprocess.cu feeds floor()/ceil() decisions -> no implicit FMA contraction
SOURCES = {
    "preprocess.cu": ["-fmad=false"],
    "scan_sort.cu": ([f"-DONESWEEP_MIN_BLOCKS={os.environ['B200GS_ONESWEEP_MIN_BLOCKS']}"]
                     if os.environ.get("B200GS_ONESWEEP_MIN_BLOCKS") else []),
    "binning.cu": [],
    "blend.cu": [],
    "loss.cu": [],
    "optim.cu": [],
    "peer.cu": [],
    "route.cu": [],
    "densify.cu": [],
    "api.cu": [],
}
HEADERS = ["common.cuh", "gs_math.cuh", os.path.join(INCLUDE, "b200gs.h")]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libb200gs.so)")


def _digest() -> str:
    h = hashlib.sha256()
    for name in list(SOURCES) + HEADERS:
        path = name if os.path.isabs(name) else os.path.join(CSRC, name)
        h.update(open(path, "rb").read())
    h.update(repr((ARCH, COMMON, SOURCES)).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(BUILD, exist_ok=True)
    stamp = os.path.join(BUILD, "stamp.txt")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for src, extra in SOURCES.items():
        obj = os.path.join(BUILD, src.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    link = [nvcc, *ARCH, "-shared", "-o", LIB, *objs]
    subprocess.run(link, check=True)
    open(stamp, "w").write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
