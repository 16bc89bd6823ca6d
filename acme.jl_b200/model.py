"""Model derivation in exact rationals (host side; runs once per model).

This is the *host* half of the reference's ``DiscreteModel`` constructor
(/root/reference/src/ACME.jl:150-262 and helpers :264-464, :717-777).  The
north_star keeps it on the Julia host; it is restated here only because no
Julia toolchain exists in this image and the tests/benchmarks need derived
model matrices for the example circuits.  Everything up to the Float64
conversion is exact (``fractions.Fraction`` <-> ``Rational{BigInt}``).

The result is a :class:`DiscreteModel` holding Float64 column-major matrices,
an element table per non-linear sub-problem and the initial solution -- i.e.
exactly what crosses the C-ABI in ``include/acmeb200.h``.
"""
from __future__ import annotations

import itertools
import warnings
from dataclasses import dataclass, field
from fractions import Fraction
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .circuit import Circuit
from .elements import NLElem, fr, reye, rzeros
from . import hostsolve

F0 = Fraction(0)
F1 = Fraction(1)


# ------------------------------------------------------------------ utilities
def consecranges(lengths: Sequence[int]) -> List[range]:
    """ACME.jl:771"""
    out, e = [], 0
    for l in lengths:
        out.append(range(e, e + l))
        e += l
    return out


def _dot(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Exact matrix product that keeps shape (and Fractions) for empty operands."""
    m, k = a.shape
    k2, n = b.shape
    assert k == k2
    out = rzeros(m, n)
    if m == 0 or n == 0 or k == 0:
        return out
    # skip zeros: the matrices are sparse and Fraction arithmetic is slow
    bnz = [[(j, b[r, j]) for j in range(n) if b[r, j] != 0] for r in range(k)]
    for i in range(m):
        row = out[i]
        for r in range(k):
            v = a[i, r]
            if v != 0:
                for j, w in bnz[r]:
                    row[j] += v * w
    return out


def _argmax_abs(m: np.ndarray) -> Tuple[int, int]:
    """``argmax(abs.(m))`` -- first maximum in column-major order."""
    best, bi, bj = None, 0, 0
    for j in range(m.shape[1]):
        for i in range(m.shape[0]):
            v = abs(m[i, j])
            if best is None or v > best:
                best, bi, bj = v, i, j
    return bi, bj


def _drop(m: np.ndarray, rows=(), cols=()) -> np.ndarray:
    r = [i for i in range(m.shape[0]) if i not in rows]
    c = [j for j in range(m.shape[1]) if j not in cols]
    return m[np.ix_(r, c)] if r and c else rzeros(len(r), len(c))


def gensolve(a: np.ndarray, b: np.ndarray, thresh=Fraction(1, 10)):
    """ACME.jl:717-747.  Returns ``(x, h)``: ``a*x == b`` and ``a*h == 0``.

    ``h`` (the running null-space basis) and ``x`` are kept column-sparse; the
    pivoting heuristics (row order by nnz, column choice by fewest non-zeros
    among those above ``thresh``) follow the reference line by line.
    """
    m, n = a.shape
    nb = b.shape[1]
    x_cols: List[Dict[int, Fraction]] = [dict() for _ in range(nb)]
    h_cols: List[Dict[int, Fraction]] = [{j: F1} for j in range(n)]
    rows_nz = [[(j, a[i, j]) for j in range(n) if a[i, j] != 0] for i in range(m)]
    order = sorted(range(m), key=lambda i: len(rows_nz[i]))  # stable, like sortperm
    for ti in order:
        ait = rows_nz[ti]
        s = []
        for hj in h_cols:
            acc = F0
            for j, v in ait:
                w = hj.get(j)
                if w is not None:
                    acc += v * w
            s.append(acc)
        nz = [(j, abs(v)) for j, v in enumerate(s) if v != 0]
        if not nz:
            continue  # numerical zero (tol is ~1e-75 for BigInt rationals)
        max_abs = max(v for _, v in nz)
        jat = [j for j, v in nz if v >= thresh * max_abs]
        j = min(jat, key=lambda jj: len(h_cols[jj]))  # first minimum
        q = h_cols[j]
        aq = s[j]
        for c in range(nb):
            xc = x_cols[c]
            acc = b[ti, c]
            for jj, v in ait:
                w = xc.get(jj)
                if w is not None:
                    acc -= v * w
            if acc != 0:
                f = acc / aq
                for r, qv in q.items():
                    nv = xc.get(r, F0) + qv * f
                    if nv == 0:
                        xc.pop(r, None)
                    else:
                        xc[r] = nv
        new_cols = []
        for jj, hj in enumerate(h_cols):
            if jj == j:
                continue
            if s[jj] != 0:
                f = s[jj] / aq
                hj = dict(hj)
                for r, qv in q.items():
                    nv = hj.get(r, F0) - qv * f
                    if nv == 0:
                        hj.pop(r, None)
                    else:
                        hj[r] = nv
            new_cols.append(hj)
        h_cols = new_cols
    x = rzeros(n, nb)
    for c, xc in enumerate(x_cols):
        for r, v in xc.items():
            x[r, c] = v
    h = rzeros(n, len(h_cols))
    for c, hc in enumerate(h_cols):
        for r, v in hc.items():
            h[r, c] = v
    return x, h


def rank_factorize(a: np.ndarray):
    """ACME.jl:749-762: ``a == c @ f`` with ``f`` of full row rank."""
    f = a
    nullspace = gensolve(a.T.copy(), rzeros(a.shape[1], 0))[1]
    c = reye(a.shape[0])
    while nullspace.shape[1] > 0:
        i, j = _argmax_abs(nullspace)
        piv = nullspace[i, j]
        ci = c[:, i].copy()
        nj = nullspace[:, j].copy()
        for col in range(c.shape[1]):
            w = nj[col] / piv
            if w != 0:
                c[:, col] = c[:, col] - ci * w
        c = _drop(c, cols=(i,))
        ni = nullspace[i, :].copy()
        for col in range(nullspace.shape[1]):
            w = ni[col] / piv
            if w != 0:
                nullspace[:, col] = nullspace[:, col] - nj * w
        nullspace = _drop(nullspace, rows=(i,), cols=(j,))
        f = _drop(f, rows=(i,))
    return c, f


# ------------------------------------------------------------------ derivation
def model_matrices(circ: Circuit, t: Fraction) -> Dict[str, np.ndarray]:
    """ACME.jl:264-315"""
    nb, nx, nq, nu = circ.nb, circ.nx, circ.nq, circ.nu
    mv, mi, mx, mxd, mq, mu = (circ.blockdiag(k) for k in ("mv", "mi", "mx", "mxd", "mq", "mu"))
    u0 = circ.u0()
    tv, ti = circ.topomat()
    nl = mv.shape[0]
    half = Fraction(1, 2)
    lhs = rzeros(nl + nb, 2 * nb + nx + nq)
    lhs[:nl, :nb] = mv
    lhs[:nl, nb:2 * nb] = mi
    if nx:
        lhs[:nl, 2 * nb:2 * nb + nx] = mxd / t + mx * half
    lhs[:nl, 2 * nb + nx:] = mq
    for r, row in enumerate(tv):
        for c_, v in enumerate(row):
            if v:
                lhs[nl + r, c_] = Fraction(v)
    for r, row in enumerate(ti):
        for c_, v in enumerate(row):
            if v:
                lhs[nl + len(tv) + r, nb + c_] = Fraction(v)
    assert len(tv) + len(ti) == nb
    rhs = rzeros(nl + nb, 1 + nu + nx)
    rhs[:nl, 0:1] = u0
    rhs[:nl, 1:1 + nu] = mu
    if nx:
        rhs[:nl, 1 + nu:] = mxd / t - mx * half
    x, f = gensolve(lhs, rhs)

    rowsizes = (nb, nb, nx, nq)
    rr = consecranges(rowsizes)
    fq = f[rr[3].start:rr[3].stop, :]
    nullspace = gensolve(fq.copy(), rzeros(fq.shape[0], 0))[1]
    indeterminates = _dot(f, nullspace)
    if sum(float(v) ** 2 for v in indeterminates[rr[2].start:rr[2].stop, :].flat) > 1e-20:
        warnings.warn("State update depends on indeterminate quantity")
    while nullspace.shape[1] > 0:
        i, j = _argmax_abs(nullspace)
        nullspace = _drop(nullspace, rows=(i,), cols=(j,))
        f = _drop(f, cols=(i,))

    def rows(m, k):
        return m[rr[k].start:rr[k].stop, :]

    mats: Dict[str, np.ndarray] = {}
    mats["fv"], mats["fi"], mats["c"], mats["fq"] = (rows(f, k).copy() for k in range(4))
    cr = consecranges((1, nu, nx))
    names = (("v0", "i0", "x0", "q0"), ("ev", "ei", "b", "eq_full"), ("dv", "di", "a", "dq_full"))
    for ci_, colr in enumerate(cr):
        for k in range(4):
            mats[names[ci_][k]] = x[rr[k].start:rr[k].stop, colr.start:colr.stop].copy()
    for v in ("v0", "i0", "x0", "q0"):
        mats[v] = mats[v][:, 0].copy()

    pv, pi, px, pxd, pq = (circ.blockdiag(k) for k in ("pv", "pi", "px", "pxd", "pq"))
    ny = pv.shape[0]
    p = rzeros(ny, 2 * nb + nx + nq)
    p[:, :nb] = pv
    p[:, nb:2 * nb] = pi
    if nx:
        p[:, 2 * nb:2 * nb + nx] = px * half + pxd / t
    p[:, 2 * nb + nx:] = pq
    if sum(float(v) ** 2 for v in _dot(p, indeterminates).flat) > 1e-20:
        warnings.warn("Model output depends on indeterminate quantity")
    dy = _dot(p, x[:, 1 + nu:])
    if nx:
        dy = dy + px * half - pxd / t
    mats["dy"] = dy
    mats["ey"] = _dot(p, x[:, 1:1 + nu])
    mats["fy"] = _dot(p, f)
    mats["y0"] = _dot(p, x[:, 0:1])[:, 0]
    return mats


def tryextract(fq: np.ndarray, numcols: int) -> Optional[np.ndarray]:
    """ACME.jl:319-347"""
    fq = fq.copy()
    n = fq.shape[1]
    a = reye(n)
    if numcols >= n:
        return a
    for colcnt in range(numcols):
        i, j = _argmax_abs(fq[:, colcnt:])
        j += colcnt
        fq[:, [colcnt, j]] = fq[:, [j, colcnt]]
        a[:, [colcnt, j]] = a[:, [j, colcnt]]
        piv = fq[i, colcnt]
        if piv == 0:
            raise ZeroDivisionError("tryextract: zero pivot")
        factors = [fq[i, jj] / piv for jj in range(colcnt + 1, n)]
        acol = a[:, colcnt].copy()
        fcol = fq[:, colcnt].copy()
        for k, jj in enumerate(range(colcnt + 1, n)):
            if factors[k] != 0:
                a[:, jj] = a[:, jj] - acol * factors[k]
                fq[:, jj] = fq[:, jj] - fcol * factors[k]
        fq = _drop(fq, rows=(i,))
        if all(v == 0 for v in fq[:, colcnt + 1:].flat):
            return a
    return None


def nldecompose(mats, nns: List[int], nqs: List[int]) -> List[List[int]]:
    """ACME.jl:349-378 (updates ``mats['fq'], mats['c'], mats['fy']`` in place)."""
    fq = mats["fq"]
    ncols = fq.shape[1]
    a = reye(ncols)
    sub_ranges = consecranges(nqs)
    extracted: List[List[int]] = []
    rem_start = 0
    rem_nles = sorted(e for e in range(len(nqs)) if nqs[e] > 0)
    while rem_nles:
        found = False
        for sz in range(1, len(rem_nles) + 1):
            for sub in itertools.combinations(rem_nles, sz):
                nn_sub = sum(nns[e] for e in sub)
                ridx = [r for e in sub for r in sub_ranges[e]]
                a_update = tryextract(fq[np.ix_(ridx, range(rem_start, ncols))], nn_sub)
                if a_update is not None:
                    fq[:, rem_start:] = _dot(fq[:, rem_start:], a_update)
                    a[:, rem_start:] = _dot(a[:, rem_start:], a_update)
                    rem_start += nn_sub
                    extracted.append(list(sub))
                    rem_nles = [e for e in rem_nles if e not in sub]
                    found = True
                    break
            if found:
                break
        if not found:
            raise RuntimeError("nldecompose: no extractable subset")
    mats["c"] = _dot(mats["c"], a)
    mats["fy"] = _dot(mats["fy"], a)
    return extracted


def split_nl_model_matrices(mats, model_qidxs, model_nns):
    """ACME.jl:381-401"""
    colr = consecranges(model_nns)
    ntot = sum(model_nns)
    fqs, fqprev_fulls = [], []
    for i, qidxs in enumerate(model_qidxs):
        block = mats["fq"][qidxs, :] if len(qidxs) else rzeros(0, ntot)
        fqs.append(block[:, colr[i].start:colr[i].stop].copy())
        prev = rzeros(len(qidxs), ntot)
        if i > 0:
            prev[:, :colr[i].start] = block[:, :colr[i].start]
        fqprev_fulls.append(prev)
    return dict(
        dq_fulls=[mats["dq_full"][qidxs, :].copy() for qidxs in model_qidxs],
        eq_fulls=[mats["eq_full"][qidxs, :].copy() for qidxs in model_qidxs],
        fqs=fqs, fqprev_fulls=fqprev_fulls,
        q0s=[mats["q0"][qidxs].copy() for qidxs in model_qidxs])


def reduce_pdims(mats, fix: bool = False):
    """ACME.jl:403-451.

    When the projection lowers the rank of ``pexp`` the component of ``pexp*p`` inside the column space of
    ``fq`` is absorbed into z: ``z = z' - P (dq x + eq u + fqprev z_prev)`` with ``P = fq_pinv * pexp``, and every
    user of z has to be corrected.  AS WRITTEN the reference corrects ``a, b, dy, ey`` and the later sub-problems'
    ``dq_full, eq_full``, but (1) its correction of the later sub-problems' ``fqprev_full[:, 1:offset]``
    (ACME.jl:435-438) is a ``copyto!`` into ``X[:, 1:offset]`` -- a fresh copy in Julia, so it has no effect -- and
    (2) ``c[:, 1:offset]`` and ``fy[:, 1:offset]`` are not corrected for the ``fqprev z_prev`` term at all.  Both only
    matter for models with several sub-problems whose reduced one has ``offset > 0`` and a non-zero ``fqprev`` (none of
    the BASELINE configs; the three-sub-problem "simplified superover" of runtests.jl:751-756 is one): there the
    as-written model deviates from the circuit equations (tests/test_independent.py measures 3 V on that circuit
    against a full-system solve).  ``fix=False`` (default) restates the reference as written, because parity is
    against the reference; ``fix=True`` applies all corrections and agrees with the full-system solve to 1e-10.
    """
    subcount = len(mats["dq_fulls"])
    dqs, eqs, fqprevs, pexps = [None] * subcount, [None] * subcount, [None] * subcount, [None] * subcount
    offset = 0
    for idx in range(subcount):
        dqf, eqf, fpf = mats["dq_fulls"][idx], mats["eq_fulls"][idx], mats["fqprev_fulls"][idx]
        pexp, dqeq = rank_factorize(np.hstack([dqf, eqf, fpf]))
        pexps[idx] = pexp
        c0, c1 = dqf.shape[1], dqf.shape[1] + eqf.shape[1]
        dqs[idx], eqs[idx], fqprevs[idx] = dqeq[:, :c0], dqeq[:, c0:c1], dqeq[:, c1:]
        fq = mats["fqs"][idx]
        nn = fq.shape[1]
        fqT = fq.T.copy()
        fq_pinv = gensolve(_dot(fqT, fq), fqT)[0]
        pexp = pexp - _dot(fq, _dot(fq_pinv, pexp))
        pexp, f = rank_factorize(pexp)
        if pexp.shape[1] < pexps[idx].shape[1]:
            cols = slice(offset, offset + nn)
            proj = _dot(fq_pinv, pexps[idx])
            cproj = _dot(mats["c"][:, cols], proj)
            fyproj = _dot(mats["fy"][:, cols], proj)
            mats["a"] = mats["a"] - _dot(cproj, dqs[idx])
            mats["b"] = mats["b"] - _dot(cproj, eqs[idx])
            mats["dy"] = mats["dy"] - _dot(fyproj, dqs[idx])
            mats["ey"] = mats["ey"] - _dot(fyproj, eqs[idx])
            if fix and offset:
                mats["c"][:, :offset] = mats["c"][:, :offset] - _dot(cproj, fqprevs[idx][:, :offset])
                mats["fy"][:, :offset] = mats["fy"][:, :offset] - _dot(fyproj, fqprevs[idx][:, :offset])
            for idx2 in range(idx + 1, subcount):
                q = _dot(mats["fqprev_fulls"][idx2][:, cols], proj)
                mats["dq_fulls"][idx2] = mats["dq_fulls"][idx2] - _dot(q, dqs[idx])
                mats["eq_fulls"][idx2] = mats["eq_fulls"][idx2] - _dot(q, eqs[idx])
                if fix and offset:  # ACME.jl:435-438 writes this into a temporary copy: no effect as written
                    mats["fqprev_fulls"][idx2][:, :offset] = (
                        mats["fqprev_fulls"][idx2][:, :offset] - _dot(q, fqprevs[idx][:, :offset]))
            pexps[idx] = pexp
            dqs[idx] = _dot(f, dqs[idx])
            eqs[idx] = _dot(f, eqs[idx])
            fqprevs[idx] = _dot(f, fqprevs[idx])
            mats["dq_fulls"][idx] = _dot(pexp, dqs[idx])
            mats["eq_fulls"][idx] = _dot(pexp, eqs[idx])
            mats["fqprev_fulls"][idx] = _dot(pexp, fqprevs[idx])
        offset += nn
    mats.update(dqs=dqs, eqs=eqs, fqprevs=fqprevs, pexps=pexps)
    return mats


def _f64(m) -> np.ndarray:
    a = np.array(m, dtype=object)
    out = np.empty(a.shape, dtype=np.float64)
    for idx, v in np.ndenumerate(a):
        out[idx] = float(v)  # correctly rounded, like Float64(::Rational{BigInt})
    return np.asfortranarray(out)


@dataclass
class SubProblem:
    """One non-linear sub-problem: matrices + element table + initial solution."""
    nn: int
    nq: int
    np_: int
    dq: np.ndarray
    eq: np.ndarray
    fqprev: np.ndarray
    pexp: np.ndarray
    q0: np.ndarray
    fq: np.ndarray
    init_z: np.ndarray
    elems: List[Tuple[NLElem, int]]  # (element, q offset)
    elem_idxs: List[int] = field(default_factory=list)


class DiscreteModel:
    """Float64 model, field for field the reference's struct (ACME.jl:118-148)."""

    def __init__(self, circ: Optional[Circuit] = None, t=None, *, decompose_nonlinearity=True,
                 solver="HomotopySolver{CachingSolver{SimpleSolver}}", fix_reduce_pdims: bool = False):
        self.solver = solver
        self.fix_reduce_pdims = fix_reduce_pdims  # see reduce_pdims: False = the reference as written
        if circ is None:
            return
        self._derive(circ, fr(t), decompose_nonlinearity)

    def __getstate__(self):
        # device runners cached by run_(model, u) hold library handles: they stay with this object
        return {k: v for k, v in self.__dict__.items() if k != "_device_runners"}

    # ACME.jl:150-262
    def _derive(self, circ: Circuit, t: Fraction, decompose: bool):
        mats = model_matrices(circ, t)
        elements = list(circ.elements.values())
        nns = [e.nn for e in elements]
        nqs = [e.nq for e in elements]
        if decompose:
            nl_elems = nldecompose(mats, nns, nqs)
        else:
            nl_elems = [[i for i, n in enumerate(nns) if n > 0]]
            if not nl_elems[0]:
                nl_elems = []
        model_nns = [sum(nns[e] for e in nles) for nles in nl_elems]
        qr = consecranges(nqs)
        model_qidxs = [[r for e in nles for r in qr[e]] for nles in nl_elems]
        mats.update(split_nl_model_matrices(mats, model_qidxs, model_nns))
        mats = reduce_pdims(mats, self.fix_reduce_pdims)
        model_nqs = [p.shape[0] for p in mats["pexps"]]
        assert circ.nn == sum(model_nns)
        tables = [circ.nl_table(nles) for nles in nl_elems]

        # initial solutions, ACME.jl:196-200 and :453-464
        init_zs = [np.zeros(n) for n in model_nns]
        fqs64 = [_f64(m) for m in mats["fqs"]]
        for idx in range(len(nl_elems)):
            zall = np.concatenate(init_zs) if init_zs else np.zeros(0)
            q = _f64(mats["q0s"][idx]) + _f64(mats["fqprev_fulls"][idx]) @ zall
            init_zs[idx] = hostsolve.initial_solution(tables[idx], fqs64[idx], q, model_nns[idx])

        # constant sub-problems, ACME.jl:202-228
        while True:
            const_idxs = [i for i, m in enumerate(mats["dqs"]) if m.shape[0] == 0]
            if not const_idxs:
                break
            zr = consecranges(model_nns)
            const_z = [r for i in const_idxs for r in zr[i]]
            varying = [r for r in range(sum(model_nns)) if r not in const_z]
            zc = np.array([fr(v) for i in const_idxs for v in init_zs[i]], dtype=object).reshape(-1, 1)
            for idx in range(len(mats["q0s"])):
                mats["q0s"][idx] = mats["q0s"][idx] + _dot(mats["fqprev_fulls"][idx][:, const_z], zc)[:, 0]
                mats["fqprev_fulls"][idx] = mats["fqprev_fulls"][idx][:, varying]
            mats["x0"] = mats["x0"] + _dot(mats["c"][:, const_z], zc)[:, 0]
            mats["y0"] = mats["y0"] + _dot(mats["fy"][:, const_z], zc)[:, 0]
            for key in ("q0s", "dq_fulls", "eq_fulls", "fqs", "fqprev_fulls"):
                mats[key] = [m for i, m in enumerate(mats[key]) if i not in const_idxs]
            keep = [i for i in range(len(init_zs)) if i not in const_idxs]
            init_zs = [init_zs[i] for i in keep]
            model_nns = [model_nns[i] for i in keep]
            model_nqs = [model_nqs[i] for i in keep]
            tables = [tables[i] for i in keep]
            nl_elems = [nl_elems[i] for i in keep]
            mats["fy"] = mats["fy"][:, varying]
            mats["c"] = mats["c"][:, varying]
            mats = reduce_pdims(mats, self.fix_reduce_pdims)

        self.a = _f64(mats["a"]); self.b = _f64(mats["b"]); self.c = _f64(mats["c"])
        self.x0 = _f64(mats["x0"])
        self.dy = _f64(mats["dy"]); self.ey = _f64(mats["ey"]); self.fy = _f64(mats["fy"])
        self.y0 = _f64(mats["y0"])
        self.subs: List[SubProblem] = []
        for idx in range(len(nl_elems)):
            self.subs.append(SubProblem(
                nn=model_nns[idx], nq=model_nqs[idx], np_=mats["dqs"][idx].shape[0],
                dq=_f64(mats["dqs"][idx]), eq=_f64(mats["eqs"][idx]), fqprev=_f64(mats["fqprevs"][idx]),
                pexp=_f64(mats["pexps"][idx]), q0=_f64(mats["q0s"][idx]), fq=_f64(mats["fqs"][idx]),
                init_z=np.array(init_zs[idx], dtype=np.float64), elems=tables[idx],
                elem_idxs=list(nl_elems[idx])))
        self.x = np.zeros(self.nx)

    @classmethod
    def from_matrices(cls, *, a, b, c, x0, dy, ey, fy, y0, subs=(), solver="HomotopySolver{CachingSolver{SimpleSolver}}"):
        """Build a model from already-derived Float64 matrices (what a Julia
        host would pass after running the reference's own derivation)."""
        m = cls(solver=solver)
        m.a, m.b, m.c = (np.asfortranarray(np.array(v, dtype=np.float64)) for v in (a, b, c))
        m.dy, m.ey, m.fy = (np.asfortranarray(np.array(v, dtype=np.float64)) for v in (dy, ey, fy))
        m.x0 = np.array(x0, dtype=np.float64).reshape(-1)
        m.y0 = np.array(y0, dtype=np.float64).reshape(-1)
        m.subs = list(subs)
        m.x = np.zeros(m.nx)
        return m

    # sizes, ACME.jl:466-472
    @property
    def nx(self): return self.x0.shape[0]
    @property
    def nu(self): return self.b.shape[1]
    @property
    def ny(self): return self.y0.shape[0]
    @property
    def nn_total(self): return sum(s.nn for s in self.subs)
    def nn(self, i=None): return self.nn_total if i is None else self.subs[i].nn
    def nq(self, i): return self.subs[i].nq
    def np(self, i): return self.subs[i].np_

    # ------------------------------------------------------------ analysis (host, off the hot path)
    def steadystate(self, u=None):
        """ACME.jl:474-497"""
        return hostsolve.steadystate(self, u)

    def linearize(self, usteady=None):
        """ACME.jl:505-550"""
        return hostsolve.linearize(self, usteady)

    def steadystate_(self, u=None):
        """``steadystate!`` (ACME.jl:499-503)"""
        xs = self.steadystate(u)
        self.x = np.array(xs, dtype=np.float64)
        return xs
