"""tests/c/abi_roundtrip.c: a plain-C host of include/acmeb200.h (what a Julia ``ccall`` binding does) -- the golden
doctest samples from the hand-derived fixture F1, create -> run -> get_solver_state -> destroy -> create ->
set_solver_state -> run == one run bit for bit, and the batch over all visible GPUs through acmeb200_multi_*.

On the GPU box it links against the product library; in the CPU suite against the host emulation of the same sources
(tests/emu), so the ABI logic (state blob, sharding, host-stream pipeline) is exercised without a device."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_roundtrip.c")


def build_and_run(libpath, tmp_path):
    exe = str(tmp_path / "abi_roundtrip")
    d, f = os.path.dirname(libpath), os.path.basename(libpath)
    subprocess.run(["gcc", "-std=c99", "-D_GNU_SOURCE", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe,
                    "-L", d, f"-l:{f}", "-lm", f"-Wl,-rpath,{d}"], check=True, capture_output=True, text=True)
    return subprocess.run([exe], capture_output=True, text=True, timeout=600)


def test_c_host_against_the_emulated_library(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    res = build_and_run(build_emu.build(), tmp_path)
    assert res.returncode == 0 and "abi_roundtrip ok" in res.stdout, res.stdout + res.stderr


@pytest.mark.gpu
def test_c_host_on_the_device(tmp_path):
    from acme_jl_b200 import _build
    res = build_and_run(_build.build() if _build.is_stale() else _build.LIB, tmp_path)
    assert res.returncode == 0 and "abi_roundtrip ok" in res.stdout, res.stdout + res.stderr
