"""ctypes front end of the CPU oracle (TEST INFRASTRUCTURE -- see acme_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  It restates the reference's
``run!``/``step!``/solver path on the CPU and is the checker for the CUDA path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libacme_oracle.so")
    src = os.path.join(_HERE, "acme_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "acmeb200.h")
    stale = (not os.path.exists(so)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(so) for f in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libacme_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.c_void_p, C.c_int64, C.c_int64]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_int]
        L.oracle_get_state.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_set_state.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.oracle_get_status.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_solve_sub.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_cache_size.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        L.oracle_cache_capacity.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        L.oracle_cache_export.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_linsolve.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_kdtree_query.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_kdtree_build.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleModel:
    """B independent copies of the reference's DiscreteModel + ModelRunner."""

    def __init__(self, model, batch=1, **desc_kw):
        from acme_jl_b200._abi import make_desc, Stats
        self._Stats = Stats
        self.model = model
        self.batch = batch
        self.holder = make_desc(model, batch, **desc_kw)
        self.h = lib().oracle_create(C.addressof(self.holder.desc), 0, batch)
        if not self.h:
            raise RuntimeError("oracle_create failed")

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_destroy(self.h)
            self.h = None

    def run(self, u, threads=1):
        """u: (nu, N) shared or (nu, N, B); returns y (ny, N, B) -- ``run!`` per instance."""
        m = self.model
        u = np.asarray(u, dtype=np.float64)
        if u.ndim == 2:
            N = u.shape[1]
            ubuf = np.ascontiguousarray(u.ravel(order="F"))
            ustride = 0
        else:
            N = u.shape[1]
            ubuf = np.ascontiguousarray(u.ravel(order="F"))
            ustride = m.nu * N
        if u.shape[0] != m.nu:
            raise ValueError(f"input matrix has {u.shape[0]} rows, but model has {m.nu} inputs")
        if ubuf.size == 0:
            ubuf = np.zeros(1)
        y = np.zeros(max(m.ny * N * self.batch, 1))
        lib().oracle_run(self.h, _p(ubuf), ustride, _p(y), m.ny * N, N, threads)
        return y[:m.ny * N * self.batch].reshape((m.ny, N, self.batch), order="F")

    @property
    def x(self):
        out = np.zeros(max(self.model.nx * self.batch, 1))
        lib().oracle_get_state(self.h, _p(out))
        return out[:self.model.nx * self.batch].reshape((self.model.nx, self.batch), order="F")

    @x.setter
    def x(self, val):
        val = np.asfortranarray(np.broadcast_to(np.asarray(val, dtype=np.float64).reshape(self.model.nx, -1),
                                                (self.model.nx, self.batch)))
        buf = np.ascontiguousarray(val.ravel(order="F"))
        lib().oracle_set_state(self.h, _p(buf), self.model.nx)

    def status(self):
        st = np.zeros(self.batch, dtype=np.uint32)
        ff = np.zeros(self.batch, dtype=np.int64)
        lib().oracle_get_status(self.h, _p(st), _p(ff))
        return st, ff

    def stats(self):
        s = self._Stats()
        lib().oracle_get_stats(self.h, C.addressof(s))
        return s.as_dict()

    def solve_sub(self, p, inst=0, sub=0):
        p = np.ascontiguousarray(p, dtype=np.float64)
        z = np.zeros(max(self.model.subs[sub].nn, 1))
        it = C.c_int(0)
        conv = lib().oracle_solve_sub(self.h, inst, sub, _p(p), _p(z), C.addressof(it))
        return z[:self.model.subs[sub].nn], bool(conv), it.value

    def export_cache(self, inst=0, sub=0):
        """Frozen copy of the learnt CachingSolver state of one instance: the
        arrays the device library takes as ``acmeb200_cache``."""
        s = self.model.subs[sub]
        cap = lib().oracle_cache_capacity(self.h, inst, sub)
        ps = np.zeros((s.np_, cap), order="F")
        zs = np.zeros((s.nn, cap), order="F")
        n = lib().oracle_cache_export(self.h, inst, sub, _p(ps), _p(zs))
        cut_dim = np.zeros(max(n - 1, 1), dtype=np.int32)
        cut_val = np.zeros(max(n - 1, 1))
        ps_idx = np.zeros(n, dtype=np.int32)
        lib().oracle_kdtree_build(s.np_, cap, n, _p(ps), _p(cut_dim), _p(cut_val), _p(ps_idx))
        return dict(cut_dim=cut_dim[:max(n - 1, 0)], cut_val=cut_val[:max(n - 1, 0)], ps_idx=ps_idx, ps=ps, zs=zs)


def linsolve(A, b):
    A = np.asfortranarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    ok = lib().oracle_linsolve(A.shape[0], _p(A), _p(b), _p(x))
    return bool(ok), x


def kdtree_query(ps, queries):
    ps = np.asfortranarray(ps, dtype=np.float64)
    queries = np.asfortranarray(queries, dtype=np.float64)
    out = np.zeros(queries.shape[1], dtype=np.int32)
    lib().oracle_kdtree_query(ps.shape[0], ps.shape[1], _p(ps), queries.shape[1], _p(queries), _p(out))
    return out
