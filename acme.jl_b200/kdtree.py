"""``KDTree`` / ``indnearest`` (/root/reference/src/kdtree.jl) through the library's host-side entry points.

The tree code is the device's own (csrc/kdcache.cuh, compiled for the host as well): ``KDTree(p, Np)`` is what a
learning ``CachingSolver`` on the device runs when it rebuilds its tree (solvers.jl:390), ``indnearest`` is its
start-point search.  A host uses them to freeze collected solutions ``(ps, zs)`` into the read-only cache a model
descriptor accepts (``caches=[...]`` of :class:`BatchRunner`)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import check, lib


class KDTree:
    """``KDTree(p, Np=size(p, 2))`` (kdtree.jl:4-73): fields ``cut_dim``, ``cut_val``, ``ps_idx`` (1-based) and ``ps``."""

    def __init__(self, p, Np=None):
        self.ps = np.asfortranarray(p, dtype=np.float64)
        if self.ps.ndim != 2:
            raise ValueError("p must be a matrix (one point per column)")
        np_, ncol = self.ps.shape
        self.n_points = ncol if Np is None else int(Np)
        self.cut_dim = np.zeros(max(self.n_points - 1, 0), dtype=np.int32)
        self.cut_val = np.zeros(max(self.n_points - 1, 0))
        self.ps_idx = np.zeros(self.n_points, dtype=np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().acmeb200_kdtree_build(np_, ncol, self.n_points, vp(self.ps), vp(self.cut_dim), vp(self.cut_val), vp(self.ps_idx)))

    def indnearest(self, p, best_dist: float = np.inf, best_pidx: int = 0):
        """1-based column of the nearest tree point for every column of ``p`` (kdtree.jl:192-234); a point must be
        strictly nearer than ``best_dist`` (squared distance), else ``best_pidx`` is returned"""
        q = np.asfortranarray(np.asarray(p, dtype=np.float64).reshape(self.ps.shape[0], -1))
        out = np.zeros(q.shape[1], dtype=np.int32)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        check(lib().acmeb200_kdtree_indnearest(self.ps.shape[0], self.ps.shape[1], self.n_points, vp(self.cut_dim), vp(self.cut_val),
                                               vp(self.ps_idx), vp(self.ps), q.shape[1], vp(q), float(best_dist), int(best_pidx), vp(out)))
        return out if np.ndim(p) > 1 else int(out[0])


def frozen_cache(ps, zs, n_points=None) -> dict:
    """the ``caches=[...]`` entry of a model descriptor (acmeb200_cache) for stored solutions ``ps`` (np x M) / ``zs``
    (nn x M): the k-d tree is built here, on the host, by the reference's constructor.  Pass the solutions alone: with
    ``n_points`` smaller than the number of columns the constructor sorts over ALL columns (kdtree.jl:37), and
    zero-filled spare columns become start points (p = 0, z = 0)"""
    t = KDTree(ps, n_points)
    zs = np.asfortranarray(zs, dtype=np.float64)
    if zs.shape[1] != t.ps.shape[1]:
        raise ValueError("ps and zs need one column per stored solution")
    return dict(cut_dim=t.cut_dim, cut_val=t.cut_val, ps_idx=t.ps_idx, ps=t.ps, zs=zs)
