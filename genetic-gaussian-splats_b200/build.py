"""Build libggs_b200.so (the C-ABI library of include/ggs_b200.h) in-tree with nvcc for sm_100a.

    python genetic-gaussian-splats_b200/build.py [--force] [--verbose]

The .so lands in genetic-gaussian-splats_b200/lib/ (git-ignored, but it travels to the GPU
box with gpurun).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
OBJ_DIR = os.path.join(HERE, "build")
LIB_PATH = os.path.join(LIB_DIR, "libggs_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"),
          "-I", CSRC]
# per-file extra flags: the decode must not contract mul+add (bit-exact AABB, SURVEY finding 4)
SOURCES = {
    "ggs_decode.cu": ["-fmad=false"],
    # the packed (64-bit) pixel accumulators live in named PTX registers (GGS_NAMED_REGS), which
    # keeps ptxas from renaming them out of place at any -O level; tests/test_cpu_sass.py guards it
    "ggs_raster.cu": ["-fmad=false"],   # it inlines the decode arithmetic (fused path)
    "ggs_breed.cu": ["-fmad=false"],    # its SA proposal kernel inlines the decode arithmetic
    "ggs_mask.cu": ["-fmad=false"],   # one rounding per operation, like the reference's torch ops
    "ggs_engine.cu": [],
    "ggs_peers.cu": [],
    "ggs_probe.cu": [],
    "ggs_api.cu": [],
}


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: cannot build libggs_b200.so")
    return exe


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "ggs_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc()] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + \
                  ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
    if force or _stale(LIB_PATH, objs):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB_PATH] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB_PATH


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
    sys.exit(0)
