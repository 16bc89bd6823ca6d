"""Loader for the CUDA library (``libacmeb200.so``, C ABI of include/acmeb200.h).

There is deliberately no CPU implementation behind this module: if the library
cannot be built/loaded, or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

from . import _build
from ._abi import ModelDesc, Stats

_LIB = None


class AcmeB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"acmeb200 error {code}: {msg}")
        self.code = code


def lib():
    global _LIB
    if _LIB is None:
        path = _build.LIB
        if _build.is_stale():
            path = _build.build()
        L = C.CDLL(path)
        vp, i64, u32 = C.c_void_p, C.c_int64, C.c_uint32
        L.acmeb200_model_create.argtypes = [C.POINTER(ModelDesc), i64, i64, C.POINTER(vp)]
        L.acmeb200_model_destroy.argtypes = [vp]
        L.acmeb200_model_destroy.restype = None
        L.acmeb200_run.argtypes = [vp, vp, i64, vp, i64, i64, u32, vp]
        L.acmeb200_get_state.argtypes = [vp, vp]
        L.acmeb200_set_state.argtypes = [vp, vp, i64]
        L.acmeb200_reset.argtypes = [vp]
        L.acmeb200_get_status.argtypes = [vp, vp, vp]
        L.acmeb200_get_stats.argtypes = [vp, C.POINTER(Stats)]
        L.acmeb200_get_cache_sizes.argtypes = [vp, C.c_int32, vp, C.POINTER(C.c_int32)]
        L.acmeb200_get_cache_info.argtypes = [vp, C.c_int32, vp]
        i32, dbl = C.c_int32, C.c_double
        L.acmeb200_kdtree_build.argtypes = [i32, i32, i32, vp, vp, vp, vp]
        L.acmeb200_kdtree_indnearest.argtypes = [i32, i32, i32, vp, vp, vp, vp, i32, vp, dbl, i32, vp]
        L.acmeb200_solver_state_size.argtypes = [vp]
        L.acmeb200_solver_state_size.restype = i64
        L.acmeb200_get_solver_state.argtypes = [vp, vp, i64]
        L.acmeb200_set_solver_state.argtypes = [vp, vp, i64]
        L.acmeb200_get_extrapolation_origin.argtypes = [vp, i32, vp, vp]
        L.acmeb200_device_count.argtypes = [C.POINTER(i32)]
        L.acmeb200_set_device.argtypes = [i32]
        L.acmeb200_get_device.argtypes = [vp, C.POINTER(i32)]
        L.acmeb200_multi_create.argtypes = [C.POINTER(ModelDesc), i64, i32, C.POINTER(vp)]
        L.acmeb200_multi_destroy.argtypes = [vp]
        L.acmeb200_multi_destroy.restype = None
        L.acmeb200_multi_shards.argtypes = [vp]
        L.acmeb200_multi_model.argtypes = [vp, i32, C.POINTER(i64), C.POINTER(i64)]
        L.acmeb200_multi_model.restype = vp
        L.acmeb200_multi_run.argtypes = [vp, vp, i64, vp, i64, i64, u32]
        L.acmeb200_diag_exp.argtypes = [vp, vp, i64]
        L.acmeb200_eval_jq.argtypes = [vp, C.c_int32, vp, vp]
        L.acmeb200_register_tpi.argtypes = [vp]
        L.acmeb200_set_kernel.argtypes = [vp, C.c_int32]
        L.acmeb200_kernel_name.argtypes = [vp]
        L.acmeb200_kernel_name.restype = C.c_char_p
        L.acmeb200_launch_count.argtypes = [vp]
        L.acmeb200_launch_count.restype = i64
        L.acmeb200_measure_fp64_peak.argtypes = [C.POINTER(C.c_double)]
        L.acmeb200_last_error.restype = C.c_char_p
        L.acmeb200_abi_version.restype = C.c_int
        _LIB = L
        if not path.endswith("_emu.so"):
            from . import _specialise  # shape plugins built earlier (acme.jl_b200/shapes/)
            _specialise.register_all()
    return _LIB


def check(rc):
    if rc != 0:
        raise AcmeB200Error(rc, lib().acmeb200_last_error().decode())


EXPORTS = ["acmeb200_model_create", "acmeb200_model_destroy", "acmeb200_run", "acmeb200_get_state",
           "acmeb200_set_state", "acmeb200_reset", "acmeb200_get_status", "acmeb200_get_stats", "acmeb200_get_cache_sizes",
           "acmeb200_get_cache_info", "acmeb200_kdtree_build", "acmeb200_kdtree_indnearest",
           "acmeb200_solver_state_size", "acmeb200_get_solver_state", "acmeb200_set_solver_state", "acmeb200_get_extrapolation_origin",
           "acmeb200_device_count", "acmeb200_set_device", "acmeb200_get_device", "acmeb200_multi_create", "acmeb200_multi_destroy",
           "acmeb200_multi_shards", "acmeb200_multi_model", "acmeb200_multi_run", "acmeb200_diag_exp", "acmeb200_eval_jq", "acmeb200_register_tpi",
           "acmeb200_set_kernel", "acmeb200_kernel_name", "acmeb200_launch_count", "acmeb200_measure_fp64_peak",
           "acmeb200_last_error", "acmeb200_abi_version"]


def measure_fp64_peak() -> float:
    """measured DFMA throughput of the current device, TFLOP/s"""
    v = C.c_double(0)
    check(lib().acmeb200_measure_fp64_peak(C.byref(v)))
    return v.value
