"""Builds lasgun_b200/liblasgun_b200.so in-tree with nvcc for sm_100a (no GPU needed)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "liblasgun_b200.so")
SOURCES = ["csrc/lgb_kernels.cu", "csrc/lgb_gpubuild.cu", "csrc/lgb_grid.cu", "csrc/lgb_api.cu", "csrc/lgb_build.cpp", "csrc/lgb_parallel.cpp", "csrc/host/lasgun_host.cpp"]
HEADERS = ["csrc/lgb_types.cuh", "csrc/lgb_gpubuild.cuh", "csrc/lgb_grid.cuh", "csrc/lgb_math.cuh", "csrc/lgb_build.hpp", "csrc/lgb_parallel.hpp", "../include/lasgun_b200.h", "../include/lasgun_host.hpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",                       # exact f64 tests must not fuse a*b+c (reference arithmetic)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "-shared",
]


def stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(os.path.join(HERE, f)) > t for f in SOURCES + HEADERS + ["build.py"])


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return SO
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", SO] + [os.path.join(HERE, f) for f in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode:
        raise RuntimeError("nvcc failed building liblasgun_b200.so")
    return SO


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(SO)
