"""Builds the CUDA library in-tree (``acme.jl_b200/libacmeb200.so``) for sm_100a.

nvcc cross-compiles without a GPU; the built .so travels to the GPU box with
the repository snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("ACMEB200_LIB") or os.path.join(HERE, "libacmeb200.so")
SOURCES = ["acmeb200.cu"]
DEPS = ["acmeb200.cu", "devmodel.h", "elements.cuh", "kernel_generic.cuh", "kernel_tpi.cuh", "kernel_coop.cuh",
        os.path.join("..", "..", "include", "acmeb200.h")]

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-cudart", "static"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    return exe


def is_stale() -> bool:
    if os.environ.get("ACMEB200_LIB"):
        return False  # explicitly chosen prebuilt variant (kernel tuning experiments)
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS if os.path.exists(os.path.join(CSRC, d)))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stdout + res.stderr)
    return LIB
