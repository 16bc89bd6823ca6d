"""Builds tests/emu/_build/libacmeb200_emu.so: the device library's OWN sources (csrc/acmeb200.cu with the
generic kernel, csrc/tpi.cu, csrc/coop.cu, csrc/rows.cu) compiled with g++ against the host emulation
of CUDA in this directory.  Test infrastructure (tests/test_emu.py); the product never loads it."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "acme.jl_b200", "csrc")
OUT = os.path.join(HERE, "_build", "libacmeb200_emu.so")
SOURCES = [os.path.join(CSRC, "acmeb200.cu"), os.path.join(CSRC, "rows.cu"), os.path.join(CSRC, "tpi.cu"), os.path.join(CSRC, "coop.cu"),
           os.path.join(HERE, "emu_runtime.cpp")]


def stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = SOURCES + [os.path.join(HERE, "cuda_runtime.h")] + \
           [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "acmeb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, defines=(), out: str = OUT) -> str:
    """`defines` (e.g. ["ACME_ROWS_SCAN_MAX=4"]) builds a variant library under another name: tests use it to push
    small cases through code paths that the product only takes at large sizes"""
    if not force and not defines and not stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    fma = ["-mfma"] if "fma" in open("/proc/cpuinfo").read() else []
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-fPIC", "-shared", "-w"] + fma + \
          [f"-D{d}" for d in defines] + ["-I", HERE, "-I", CSRC, "-o", out]
    for s in SOURCES:
        cmd += ["-x", "c++", s]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + (res.stdout + res.stderr)[-6000:])
    return out


if __name__ == "__main__":
    print(build(force=True))
