"""ctypes mirror of ``include/acmeb200.h`` and the descriptor builder.

:func:`make_desc` turns a :class:`DiscreteModel` plus the per-instance sweeps
(element parameters, matrices, initial solutions) into the
``acmeb200_model_desc`` that crosses the C ABI.  The same structs are consumed
by the CPU oracle (tests only), which is why they live here.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

ABI_VERSION = 1
HIST_BINS = 32

SOLVER_SIMPLE = 0
SOLVER_HOMOTOPY = 1
SOLVER_HOMOTOPY_CACHING = 2
SOLVERS = {
    "SimpleSolver": SOLVER_SIMPLE,
    "HomotopySolver{SimpleSolver}": SOLVER_HOMOTOPY,
    "HomotopySolver{CachingSolver{SimpleSolver}}": SOLVER_HOMOTOPY_CACHING,
}

STATUS_NOT_CONVERGED = 1
STATUS_NONFINITE = 2
U_DEVICE = 1
Y_DEVICE = 2
SAMPLE_MAJOR = 4  # (nu, B, N) / (ny, B, N) streams, strides = sample pitches


class Array(C.Structure):
    _fields_ = [("ptr", C.POINTER(C.c_double)), ("stride", C.c_int64)]


class Elem(C.Structure):
    _fields_ = [("kind", C.c_int32), ("q_offset", C.c_int32),
                ("param_offset", C.c_int32), ("nparam", C.c_int32)]


class Cache(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("n_columns", C.c_int32),
                ("cut_dim", C.POINTER(C.c_int32)), ("cut_val", C.POINTER(C.c_double)),
                ("ps_idx", C.POINTER(C.c_int32)), ("ps", C.POINTER(C.c_double)),
                ("zs", C.POINTER(C.c_double))]


class SubDesc(C.Structure):
    _fields_ = [("nn", C.c_int32), ("nq", C.c_int32), ("np", C.c_int32), ("nelem", C.c_int32),
                ("dq", Array), ("eq", Array), ("fqprev", Array), ("pexp", Array),
                ("q0", Array), ("fq", Array), ("init_z", Array),
                ("elems", C.POINTER(Elem)), ("params", Array),
                ("nparams", C.c_int32), ("reserved", C.c_int32), ("cache", Cache)]


class ModelDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("nx", C.c_int32), ("nu", C.c_int32),
                ("ny", C.c_int32), ("nsub", C.c_int32), ("solver", C.c_int32),
                ("maxiter", C.c_int32), ("cache_capacity", C.c_int32), ("tol", C.c_double),
                ("a", Array), ("b", Array), ("c", Array), ("x0", Array),
                ("dy", Array), ("ey", Array), ("fy", Array), ("y0", Array),
                ("subs", C.POINTER(SubDesc))]


class Stats(C.Structure):
    _fields_ = [("samples", C.c_uint64), ("solves", C.c_uint64), ("newton_iters", C.c_uint64),
                ("homotopy_solves", C.c_uint64), ("not_converged", C.c_uint64),
                ("iter_hist", C.c_uint64 * HIST_BINS)]

    def as_dict(self):
        return dict(samples=self.samples, solves=self.solves, newton_iters=self.newton_iters,
                    homotopy_solves=self.homotopy_solves, not_converged=self.not_converged,
                    iter_hist=list(self.iter_hist))


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class DescHolder:
    """Owns the numpy buffers a descriptor points into."""

    def __init__(self):
        self.keep: List[object] = []
        self.desc: Optional[ModelDesc] = None

    def array(self, value, shape, batch: int) -> Array:
        """``value`` has ``shape`` (shared) or ``shape + (batch,)`` (per instance),
        column-major either way, as a Julia ``Array{Float64}`` would be."""
        a = np.asarray(value, dtype=np.float64)
        shape = tuple(int(v) for v in shape)
        n = int(np.prod(shape)) if len(shape) else 1
        if a.size == n:
            if a.shape != shape:
                a = a.reshape(shape, order="F")
            buf = np.ascontiguousarray(a.ravel(order="F"))
            stride = 0
        elif a.size == n * batch:
            if a.shape != shape + (batch,):
                raise ValueError(f"per-instance array must have shape {shape + (batch,)}, got {a.shape}")
            buf = np.ascontiguousarray(a.ravel(order="F"))
            stride = n
        else:
            raise ValueError(f"array of size {a.size} matches neither {shape} nor {shape}+({batch},)")
        if buf.size == 0:
            buf = np.zeros(1)
        self.keep.append(buf)
        return Array(_dptr(buf), stride)


def make_desc(model, batch: int = 1, *, params: Optional[Sequence[Optional[np.ndarray]]] = None,
              overrides: Optional[Dict[str, np.ndarray]] = None,
              init_z: Optional[Sequence[Optional[np.ndarray]]] = None,
              caches: Optional[Sequence[Optional[dict]]] = None,
              solver: Optional[str] = None, tol: float = 0.0, maxiter: int = 0, cache_capacity: int = 0) -> DescHolder:
    """Build the C descriptor.

    params[i]     (nparams_i, batch) element parameters of sub i (None -> shared, from the model)
    overrides     per-instance linear matrices: keys a,b,c,x0,dy,ey,fy,y0 (shape + (batch,)) and
                  ``dq{i}``, ``eq{i}``, ``fqprev{i}``, ``pexp{i}``, ``q0{i}``, ``fq{i}`` for sub i
    init_z[i]     (nn_i, batch) per-instance initial solutions (None -> the model's)
    caches[i]     dict(cut_dim, cut_val, ps_idx, ps, zs) frozen k-d tree for sub i (see acme_jl_b200.kdtree.frozen_cache)
    cache_capacity  stored solutions per instance the learning CachingSolver can hold (0: automatic)
    """
    h = DescHolder()
    ov = overrides or {}
    nx, nu, ny = model.nx, model.nu, model.ny
    nnt = model.nn_total
    nsub = len(model.subs)
    subs = (SubDesc * max(nsub, 1))()
    h.keep.append(subs)
    for i, s in enumerate(model.subs):
        sd = subs[i]
        sd.nn, sd.nq, sd.np, sd.nelem = s.nn, s.nq, s.np_, len(s.elems)
        sd.dq = h.array(ov.get(f"dq{i}", s.dq), (s.np_, nx), batch)
        sd.eq = h.array(ov.get(f"eq{i}", s.eq), (s.np_, nu), batch)
        sd.fqprev = h.array(ov.get(f"fqprev{i}", s.fqprev), (s.np_, nnt), batch)
        sd.pexp = h.array(ov.get(f"pexp{i}", s.pexp), (s.nq, s.np_), batch)
        sd.q0 = h.array(ov.get(f"q0{i}", s.q0), (s.nq,), batch)
        sd.fq = h.array(ov.get(f"fq{i}", s.fq), (s.nq, s.nn), batch)
        iz = init_z[i] if init_z is not None and init_z[i] is not None else s.init_z
        sd.init_z = h.array(iz, (s.nn,), batch)
        elems = (Elem * max(len(s.elems), 1))()
        h.keep.append(elems)
        flat: List[float] = []
        for k, (e, qoff) in enumerate(s.elems):
            elems[k] = Elem(e.kind, qoff, len(flat), len(e.params))
            flat.extend(e.params)
        sd.elems = elems
        sd.nparams = len(flat)
        pv = params[i] if params is not None and params[i] is not None else np.array(flat, dtype=np.float64)
        sd.params = h.array(pv, (len(flat),), batch)
        cache = caches[i] if caches is not None and caches[i] is not None else None
        if cache is not None:
            cd = np.ascontiguousarray(cache["cut_dim"], dtype=np.int32)
            cv = np.ascontiguousarray(cache["cut_val"], dtype=np.float64)
            pi = np.ascontiguousarray(cache["ps_idx"], dtype=np.int32)
            ps = np.asfortranarray(cache["ps"], dtype=np.float64)
            zs = np.asfortranarray(cache["zs"], dtype=np.float64)
            if cd.size == 0:
                cd = np.zeros(1, dtype=np.int32)
                cv = np.zeros(1)
            h.keep += [cd, cv, pi, ps, zs]
            sd.cache = Cache(len(pi), ps.shape[1], cd.ctypes.data_as(C.POINTER(C.c_int32)), _dptr(cv),
                             pi.ctypes.data_as(C.POINTER(C.c_int32)), _dptr(ps), _dptr(zs))
    d = ModelDesc()
    d.abi_version = ABI_VERSION
    d.nx, d.nu, d.ny, d.nsub = nx, nu, ny, nsub
    d.solver = SOLVERS[solver or model.solver]
    d.maxiter = maxiter
    d.cache_capacity = cache_capacity
    d.tol = tol
    d.a = h.array(ov.get("a", model.a), (nx, nx), batch)
    d.b = h.array(ov.get("b", model.b), (nx, nu), batch)
    d.c = h.array(ov.get("c", model.c), (nx, nnt), batch)
    d.x0 = h.array(ov.get("x0", model.x0), (nx,), batch)
    d.dy = h.array(ov.get("dy", model.dy), (ny, nx), batch)
    d.ey = h.array(ov.get("ey", model.ey), (ny, nu), batch)
    d.fy = h.array(ov.get("fy", model.fy), (ny, nnt), batch)
    d.y0 = h.array(ov.get("y0", model.y0), (ny,), batch)
    d.subs = subs
    h.desc = d
    return h
