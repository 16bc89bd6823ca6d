"""Compile-time specialisation for shapes the library does not carry.

The thread-per-instance kernel (csrc/kernel_tpi.cuh) is a template over the model's dimensions and its sequence of
non-linear elements; ``libacmeb200.so`` instantiates it for the BASELINE circuits.  Any other single-sub-problem model
with a small non-linear system otherwise runs on the run-time-dimension kernels (2-100x slower).  ``specialise(model)``
generates the one-line instantiation for the model's shape (csrc/shape_plugin.cu.in), builds it with nvcc for sm_100a
into ``acme.jl_b200/shapes/libacmeb200_shape_<key>.so`` and registers it with the loaded library
(``acmeb200_register_tpi``): from then on ``BatchRunner`` picks the specialised kernel for every model of that shape.
Plugins found in ``shapes/`` are registered when the library is loaded, so a shape is built once.

A Julia host does the same with the template file and one nvcc call; nothing here needs Python at run time."""
from __future__ import annotations

import ctypes as C
import glob
import os
import subprocess

from . import _build

HERE = os.path.dirname(os.path.abspath(__file__))
SHAPES = os.path.join(HERE, "shapes")
ELEM_TYPES = {1: "Diode", 2: "Bjt", 3: "Pot", 4: "Mosfet", 5: "OpampTanh", 6: "JilesAtherton", 100: "TestQuad"}
MAX_NN, MAX_NX, MAX_NP = 4, 8, 8   # one thread holds the Jacobian, its LU and the state in registers
_loaded = {}


def shape_of(model):
    """(nx, nu, ny, np, element kinds) of a model the thread-per-instance kernel can be instantiated for"""
    if len(model.subs) > 1:
        raise ValueError("models with several non-linear sub-problems run on the lanes-per-instance / generic kernels")
    if not model.subs:
        return (model.nx, model.nu, model.ny, 0, ())
    s = model.subs[0]
    kinds = tuple(e.kind for e, _ in s.elems)
    if s.nn > MAX_NN or model.nx > MAX_NX or s.np_ > MAX_NP:
        raise ValueError(f"nn = {s.nn}, nx = {model.nx}, np = {s.np_}: too large for one thread's registers (limits {MAX_NN}, {MAX_NX}, {MAX_NP}); "
                         "such systems run on the warp-per-instance / lanes-per-instance kernels")
    if any(k not in ELEM_TYPES for k in kinds):
        raise ValueError("unknown element kind")
    return (model.nx, model.nu, model.ny, s.np_, kinds)


def shape_key(shape) -> str:
    nx, nu, ny, np_, kinds = shape
    return f"nx{nx}_nu{nu}_ny{ny}_np{np_}_" + ("e" + "_".join(str(k) for k in kinds) if kinds else "linear")


def plugin_source(shape) -> str:
    nx, nu, ny, np_, kinds = shape
    name = f"tpi<specialised nx{nx} nu{nu} ny{ny} np{np_} [{','.join(ELEM_TYPES[k].lower() for k in kinds)}]>"
    src = open(os.path.join(_build.CSRC, "shape_plugin.cu.in")).read()
    for k, v in (("@NX@", nx), ("@NU@", nu), ("@NY@", ny), ("@NP@", np_), ("@NAME@", name),
                 ("@ELEMS@", "".join(", " + ELEM_TYPES[k] for k in kinds))):
        src = src.replace(k, str(v))
    return src


def build_plugin(shape, force: bool = False) -> str:
    os.makedirs(SHAPES, exist_ok=True)
    key = shape_key(shape)
    so, cu = os.path.join(SHAPES, f"libacmeb200_shape_{key}.so"), os.path.join(SHAPES, f"shape_{key}.cu")
    deps = [os.path.join(_build.CSRC, f) for f in os.listdir(_build.CSRC)]
    if not force and os.path.exists(so) and all(os.path.getmtime(d) <= os.path.getmtime(so) for d in deps):
        return so
    open(cu, "w").write(plugin_source(shape))
    cmd = [_build.nvcc()] + _build.NVCC_FLAGS + ["-I", _build.CSRC, "-shared", "-o", so, cu]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed on the shape plugin:\n" + res.stdout + res.stderr)
    return so


def register_plugin(path: str) -> None:
    from ._lib import check, lib
    if path in _loaded:
        return
    plug = C.CDLL(path, mode=C.RTLD_GLOBAL)
    plug.acmeb200_shape_entry.restype = C.c_void_p
    check(lib().acmeb200_register_tpi(C.c_void_p(plug.acmeb200_shape_entry())))
    _loaded[path] = plug


def register_all() -> None:
    """plugins built earlier (called when the library is loaded)"""
    for path in sorted(glob.glob(os.path.join(SHAPES, "libacmeb200_shape_*.so"))):
        try:
            register_plugin(path)
        except Exception:
            pass  # a stale plugin (built against other kernel sources) is rebuilt by the next specialise() call


def specialise(model, force: bool = False) -> str:
    """build (once) and register the thread-per-instance kernel for this model's shape; returns the kernel's name prefix"""
    if os.environ.get("ACMEB200_LIB", "").endswith("_emu.so"):
        raise RuntimeError("the host emulation cannot load CUDA shape plugins")
    shape = shape_of(model)
    register_plugin(build_plugin(shape, force))
    return shape_key(shape)
