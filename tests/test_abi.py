"""CPU-only checks of the C-ABI boundary: the library builds for sm_100a, loads,
exports every symbol include/acmeb200.h declares, the ctypes mirror matches the
header's struct sizes, and the product path refuses to run without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import acme_jl_b200 as A
from acme_jl_b200 import _abi, _build, examples as ex
from acme_jl_b200._lib import EXPORTS, AcmeB200Error, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "acmeb200.h")).read()
    declared = set(re.findall(r"\b(acmeb200_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(EXPORTS)
    L = lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.acmeb200_abi_version() == _abi.ABI_VERSION


def test_struct_sizes_match_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "acmeb200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(acmeb200_array),sizeof(acmeb200_elem),sizeof(acmeb200_cache),sizeof(acmeb200_sub_desc),'
                   'sizeof(acmeb200_model_desc),sizeof(acmeb200_stats));return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_abi.Array), C.sizeof(_abi.Elem), C.sizeof(_abi.Cache), C.sizeof(_abi.SubDesc),
                     C.sizeof(_abi.ModelDesc), C.sizeof(_abi.Stats)]


def test_library_is_sm100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", _build.build()], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    with pytest.raises(AcmeB200Error) as ei:
        A.run_(ex.diodeclipper(), np.zeros((1, 8)))
    assert ei.value.code == -5


def test_bad_descriptor_rejected():
    h = _abi.make_desc(ex.diodeclipper(), 1)
    h.desc.abi_version = 99
    out = C.c_void_p()
    assert lib().acmeb200_model_create(C.byref(h.desc), 0, 1, C.byref(out)) == -1
    assert b"ABI version" in lib().acmeb200_last_error()


def test_null_pointers_in_descriptor_rejected():
    """a descriptor with counts but no arrays is refused before anything touches the device"""
    out = C.c_void_p()

    def create(h):
        return lib().acmeb200_model_create(C.byref(h.desc), 0, 1, C.byref(out))

    h = _abi.make_desc(ex.diodeclipper(), 1)
    h.desc.subs[0].elems = None
    assert create(h) == -1 and b"elems is null" in lib().acmeb200_last_error()
    h = _abi.make_desc(ex.diodeclipper(), 1)
    h.desc.subs[0].params.ptr = None
    assert create(h) == -1 and b"params is null" in lib().acmeb200_last_error()
    h = _abi.make_desc(ex.diodeclipper(), 1)
    h.desc.subs[0].cache.n_points = 4          # a frozen cache without its arrays
    assert create(h) == -1 and b"incomplete frozen cache" in lib().acmeb200_last_error()
    h = _abi.make_desc(ex.diodeclipper(), 1)
    h.desc.subs = None
    assert create(h) == -1 and b"subs is null" in lib().acmeb200_last_error()
    assert not out.value
