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
SOURCES = ["acmeb200.cu", "tpi.cu", "tpi_wide.cu", "coop.cu", "rows.cu"]  # one translation unit per kernel family
HEADERS = ["devmodel.h", "hostmodel.h", "tma.cuh", "elements.cuh", "kernel_generic.cuh", "kdcache.cuh", "kdcache_warp.cuh", "tpi_launch.cuh", "tpi_shapes.h", "kernel_tpi.cuh",
           "kernel_coop.cuh", "kernel_rows.cuh", os.path.join("..", "..", "include", "acmeb200.h")]
# what each translation unit includes (besides devmodel.h / hostmodel.h / acmeb200.h, which all do)
TU_DEPS = {
    "acmeb200.cu": ["elements.cuh", "kernel_generic.cuh"],
    "tpi.cu": ["elements.cuh", "kernel_generic.cuh", "kernel_tpi.cuh", "tma.cuh", "tpi_launch.cuh", "tpi_shapes.h", "kdcache_warp.cuh"],
    "tpi_wide.cu": ["elements.cuh", "kernel_generic.cuh", "kernel_tpi.cuh", "tma.cuh", "tpi_launch.cuh", "tpi_shapes.h", "kdcache_warp.cuh"],
    "coop.cu": ["elements.cuh", "kernel_generic.cuh", "kernel_coop.cuh", "tma.cuh"],
    "rows.cu": ["elements.cuh", "kernel_generic.cuh", "kernel_coop.cuh", "kernel_rows.cuh", "tma.cuh", "kdcache_warp.cuh"],
}
COMMON_DEPS = ["devmodel.h", "hostmodel.h", "kdcache.cuh", os.path.join("..", "..", "include", "acmeb200.h")]
OBJDIR = os.path.join(HERE, "build")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-cudart", "static"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")
    return exe


def _mtime(rel: str) -> float:
    p = os.path.join(CSRC, rel)
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _obj(src: str) -> str:
    return os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")


def _tu_stale(src: str) -> bool:
    o = _obj(src)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(_mtime(d) > t for d in [src] + TU_DEPS[src] + COMMON_DEPS)


def is_stale() -> bool:
    if os.environ.get("ACMEB200_LIB"):
        return False  # explicitly chosen prebuilt variant (kernel tuning experiments)
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(_mtime(d) > t for d in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, extra_flags=()) -> str:
    """Compile the stale translation units in parallel (sm_100a), then link the shared library."""
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    todo = [s for s in SOURCES if force or _tu_stale(s)]
    procs = []
    for src in todo:
        cmd = [nvcc()] + NVCC_FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + \
              ["-c", "-o", _obj(src), os.path.join(CSRC, src)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out = p.communicate()[0]
        log.append(out)
        if p.returncode != 0:
            for _, q in procs:
                if q.poll() is None:
                    q.kill()
            raise RuntimeError(f"nvcc failed on {src}:\n" + out)
    cmd = [nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
           "-o", LIB] + [_obj(s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if verbose:
        print("\n".join(log))
    return LIB
