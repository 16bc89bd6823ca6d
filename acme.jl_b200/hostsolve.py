"""Derivation-time solves on the host (NOT the run! hot path).

The reference finds the initial operating point of every non-linear
sub-problem while *constructing* a ``DiscreteModel``
(``initial_solution``, /root/reference/src/ACME.jl:453-464) and offers
``steadystate`` (:474-497); both use a ``HomotopySolver{SimpleSolver}`` on the
host.  The north_star leaves model derivation on the host, so this small numpy
Newton/homotopy lives here.  It is never used by ``run_``/``BatchRunner``: the
per-sample path exists only as CUDA (see ``csrc/``) and fails loudly without it.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

from .elements import (KIND_BJT, KIND_DIODE, KIND_JA, KIND_MOSFET, KIND_OPAMP_TANH,
                       KIND_POT, NLElem)


def _evalpoly(x, coeffs):
    acc = coeffs[-1]
    for c in reversed(coeffs[:-1]):
        acc = x * acc + c
    return acc


def eval_element(e: NLElem, q: Sequence[float]):
    """Element laws, elements.jl:25-30, 107-129, 238-244, 323-401, 453-479, 540-546.
    Returns ``(res[nn], J[nn][nq])``."""
    P = e.params
    inf = math.inf
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        if e.kind == KIND_DIODE:
            is_, eta = P
            v, i = q
            ex = np.exp(np.float64(v) * (1 / (25e-3 * eta)))
            return [is_ * (ex - 1) - i], [[is_ / (25e-3 * eta) * ex, -1.0]]
        if e.kind == KIND_POT:
            (r,) = P
            v1, v2, i1, i2, pos = q
            return ([v1 - r * pos * i1, v2 - r * (1 - pos) * i2],
                    [[1, 0, -r * pos, 0, -r * i1], [0, 1, 0, -r * (1 - pos), -r * i2]])
        if e.kind == KIND_OPAMP_TANH:
            gain, scale = P
            vi, vo = q
            vs = np.float64(vi) * (gain / scale)
            return [np.tanh(vs) * scale - vo], [[gain / np.cosh(vs) ** 2, -1.0]]
        if e.kind == KIND_BJT:
            ise, isc, ηe, ηc, βf, βr, ile, ilc, ηel, ηcl, vaf, var, ikf, ikr = P
            vE, vC, iE, iC = (np.float64(v) for v in q)
            expE = np.exp(vE * (1 / (25e-3 * ηe)))
            expC = np.exp(vC * (1 / (25e-3 * ηc)))
            i_f = (βf / (1 + βf) * ise) * (expE - 1)
            i_r = (βr / (1 + βr) * isc) * (expC - 1)
            di_f1 = (βf / (1 + βf) * ise / (25e-3 * ηe)) * expE
            di_r2 = (βr / (1 + βr) * isc / (25e-3 * ηc)) * expC
            early = not (var == inf and vaf == inf)
            knee = not (ikf == inf and ikr == inf)
            if not early and not knee:
                i_cc, di_cc1, di_cc2 = i_f - i_r, di_f1, -di_r2
            elif early and not knee:
                q1 = 1 - vE * (1 / var) - vC * (1 / vaf)
                i_cc = q1 * (i_f - i_r)
                di_cc1 = (-1 / var) * (i_f - i_r) + q1 * di_f1
                di_cc2 = (-1 / vaf) * (i_f - i_r) - q1 * di_r2
            elif not early and knee:
                q2 = i_f * (1 / ikf) + i_r * (1 / ikr)
                qden = 1 + np.sqrt(1 + 4 * q2)
                qfact = 2 / qden
                i_cc = qfact * (i_f - i_r)
                dq21 = di_f1 * (1 / ikf)
                dq22 = di_r2 * (1 / ikr)
                dqf1 = -4 * dq21 / (qden - 1) / qden ** 2
                dqf2 = -4 * dq22 / (qden - 1) / qden ** 2
                di_cc1 = dqf1 * (i_f - i_r) + qfact * di_f1
                di_cc2 = dqf2 * (i_f - i_r) - qfact * di_r2
            else:
                q1 = 1 - vE * (1 / var) - vC * (1 / vaf)
                q2 = i_f * (1 / ikf) + i_r * (1 / ikr)
                qden = 1 + np.sqrt(1 + 4 * q2)
                qfact = 2 * q1 / qden
                i_cc = qfact * (i_f - i_r)
                dq11, dq12 = -1 / var, -1 / vaf
                dq21 = di_f1 * (1 / ikf)
                dq22 = di_r2 * (1 / ikr)
                dqf1 = (2 * dq11 * qden - q1 * 4 * dq21 / (qden - 1)) / qden ** 2
                dqf2 = (2 * dq12 * qden - q1 * 4 * dq22 / (qden - 1)) / qden ** 2
                di_cc1 = dqf1 * (i_f - i_r) + qfact * di_f1
                di_cc2 = dqf2 * (i_f - i_r) - qfact * di_r2
            iBE = (1 / βf) * i_f
            diBE1 = (1 / βf) * di_f1
            if ile != 0:
                expEl = np.exp(vE * (1 / (25e-3 * ηel))) if ηel != ηe else expE
                iBE += ile * (expEl - 1)
                diBE1 += (ile / (25e-3 * ηe)) * expEl
            iBC = (1 / βr) * i_r
            diBC2 = (1 / βr) * di_r2
            if ilc != 0:
                expCl = np.exp(vC * (1 / (25e-3 * ηcl))) if ηcl != ηc else expC
                iBC += ilc * (expCl - 1)
                diBC2 += (ilc / (25e-3 * ηc)) * expCl
            return ([i_cc + iBE - iE, -i_cc + iBC - iC],
                    [[di_cc1 + diBE1, di_cc2, -1.0, 0.0], [-di_cc1, -di_cc2 + diBC2, 0.0, -1.0]])
        if e.kind == KIND_MOSFET:
            pol, lam = P[0], P[1]
            nvt, nal = int(P[2]), int(P[3])
            vt = P[4:4 + nvt]
            al = P[8:8 + nal]
            dvt = [vt[k] * k for k in range(1, nvt)]
            dal = [al[k] * k for k in range(1, nal)]
            vgs, vds, id_ = q
            a_ = _evalpoly(pol * vgs, al)
            da = _evalpoly(pol * vgs, dal) if dal else 0
            vt_ = _evalpoly(pol * vgs, vt)
            dvt_ = _evalpoly(pol * vgs, dvt) if dvt else 0
            lam_ = lam if vds >= 0 else 0.0
            if vgs <= vt_:
                return [-id_], [[0.0, 0.0, -1.0]]
            if vds <= vgs - vt_:
                return ([a_ * (vgs - vt_ - 0.5 * vds) * vds * (1 + lam_ * vds) - id_],
                        [[a_ * (1 - dvt_) * vds * (1 + lam_ * vds)
                          + da * (vgs - vt_ - 0.5 * vds) * vds * (1 + lam_ * vds),
                          a_ * (vgs - vt_ + vds * (2 * lam_ * (vgs - vt_ - 0.75 * vds) - 1)), -1.0]])
            return ([(a_ / 2) * (vgs - vt_) ** 2 * (1 + lam_ * vds) - id_],
                    [[a_ * (vgs - vt_) * (1 - dvt_) * (1 + lam_ * vds)
                      + da / 2 * (vgs - vt_) ** 2 * (1 + lam_ * vds),
                      lam_ * a_ / 2 * (vgs - vt_) ** 2, -1.0]])
        if e.kind == KIND_JA:
            Ms, a, α, c, k = P
            q1, q2, q3, q4 = (np.float64(v) for v in q)
            coth = 1 / np.tanh(q1)
            aq = abs(q1)
            L = q1 / 3 if aq < 1e-4 else coth - 1 / q1
            Ld = 1 / 3 if aq < 1e-4 else 1 / q1 ** 2 - coth ** 2 + 1
            Ld2 = -2 / 15 * q1 if aq < 1e-3 else 2 * coth * (coth ** 2 - 1) - 2 / q1 ** 3
            δ = 1.0 if q3 > 0 else -1.0
            Man = Ms * L
            δM = 1.0 if np.sign(q3) == np.sign(Man - q2) else 0.0
            den = δ * (k * (1 - c)) - α * (Man - q2)
            s = 1e-4 / Ms
            res = s * ((1 - c) * δM * (Man - q2) / den * q3 + (c * Ms / a) * (q3 + α * q4) * Ld - q4)
            J1 = s * (((1 - c) ** 2 * k * Ms) * δM * Ld * δ / den ** 2 * q3 + (c * Ms / a) * (q3 + α * q4) * Ld2)
            J2 = s * -(1 - c) ** 2 * k * δM * δ / den ** 2 * q3
            J3 = s * ((1 - c) * δM * (Man - q2) / den + (c * Ms / a) * Ld)
            J4 = s * ((c * Ms / a * α) * Ld - 1)
            return [res], [[J1, J2, J3, J4]]
    raise ValueError(f"unknown element kind {e.kind}")


def eval_table(table: List[Tuple[NLElem, int]], q: np.ndarray, nn: int):
    """``CircuitNLFunc`` (circuit.jl:6-20): concatenated residuals, block-diagonal Jq."""
    res = np.zeros(nn)
    Jq = np.zeros((nn, len(q)))
    row = 0
    for e, off in table:
        r, J = eval_element(e, q[off:off + e.nq])
        for k in range(e.nn):
            res[row + k] = r[k]
            Jq[row + k, off:off + e.nq] = J[k]
        row += e.nn
    return res, Jq


class _Simple:
    """SimpleSolver (solvers.jl:151-236) for ``res(q0' + pexp*p + fq*z)``."""

    def __init__(self, table, fq, q0, pexp, nn, p0, z0, tol=1e-10):
        self.table, self.fq, self.q0, self.pexp, self.nn, self.tol = table, fq, q0, pexp, nn, tol
        self.iters = 0
        self.resmaxabs = 0.0
        self.set_origin(np.array(p0, dtype=float), np.array(z0, dtype=float))

    def _eval(self, p, z):
        q = self.q0 + self.pexp @ p + self.fq @ z
        res, Jq = eval_table(self.table, q, self.nn)
        return res, Jq @ self.fq, Jq

    def set_origin(self, p, z):
        _, J, Jq = self._eval(p, z)
        self.last_J, self.last_Jp, self.last_p, self.last_z = J, Jq @ self.pexp, p.copy(), z.copy()

    def converged(self):
        return self.resmaxabs < self.tol

    def solve(self, p, maxiter=500):
        with np.errstate(all="ignore"):
            try:
                z = self.last_z - np.linalg.solve(self.last_J, self.last_Jp @ (p - self.last_p)) \
                    if self.nn else np.zeros(0)
            except np.linalg.LinAlgError:
                z = self.last_z.copy()
            for self.iters in range(1, maxiter + 1):
                res, J, Jq = self._eval(p, z)
                self.resmaxabs = float(np.max(np.abs(res))) if self.nn else 0.0
                if not math.isfinite(self.resmaxabs) or not np.all(np.isfinite(J)):
                    return z
                if self.converged():
                    break
                try:
                    z = z - np.linalg.solve(J, res)
                except np.linalg.LinAlgError:
                    return z
            if self.converged():
                self.last_J, self.last_Jp, self.last_p, self.last_z = J, Jq @ self.pexp, p.copy(), z.copy()
        return z


def homotopy_solve(base: _Simple, p):
    """HomotopySolver.solve (solvers.jl:268-296)."""
    z = base.solve(p)
    if not base.converged():
        a, best_a = 0.5, 0.0
        start_p = base.last_p.copy()
        while best_a < 1:
            pa = (1 - a) * start_p + a * p
            z = base.solve(pa)
            if base.converged():
                best_a, a = a, 1.0
            else:
                new_a = (a + best_a) / 2
                if not (best_a < new_a < a):
                    break
                a = new_a
    return z


def initial_solution(table, fq, q0, nn):
    """ACME.jl:453-464: homotopy in q0 from 0 to the true q0."""
    nq = len(q0)
    base = _Simple(table, fq, np.zeros(nq), np.eye(nq), nn, np.zeros(nq), np.zeros(nn))
    z = homotopy_solve(base, np.array(q0, dtype=float))
    if not base.converged():
        raise RuntimeError("Failed to find initial solution")
    return z


def steadystate(model, u=None):
    """ACME.jl:474-497"""
    u = np.zeros(model.nu) if u is None else np.array(u, dtype=float)
    IA = np.eye(model.nx) - model.a
    solveIA = (lambda m: np.linalg.solve(IA, m)) if model.nx else (lambda m: m)
    steady_z = np.zeros(model.nn_total)
    zoff = 0
    for s in model.subs:
        dqIA = s.dq @ np.linalg.inv(IA) if model.nx else s.dq
        steady_q0 = s.q0 + s.pexp @ ((dqIA @ model.b + s.eq) @ u + (dqIA @ model.c + s.fqprev) @ steady_z) \
            + s.pexp @ dqIA @ model.x0
        fq = s.pexp @ dqIA @ model.c[:, zoff:zoff + s.nn] + s.fq
        base = _Simple(s.elems, fq, np.zeros(s.nq), np.eye(s.nq), s.nn, np.zeros(s.nq), np.zeros(s.nn), tol=1e-15)
        z = homotopy_solve(base, steady_q0)
        if not base.converged():
            raise RuntimeError("Failed to find steady state solution")
        steady_z[zoff:zoff + s.nn] = z
        zoff += s.nn
    return solveIA(model.b @ u + model.c @ steady_z + model.x0)


def linearize(model, usteady=None):
    """``linearize(model, usteady)`` (ACME.jl:505-550, solvers.jl:407-414): small-signal linear
    ``DiscreteModel`` around the steady state for the constant input ``usteady``."""
    from .model import DiscreteModel
    usteady = np.zeros(model.nu) if usteady is None else np.array(usteady, dtype=float)
    xsteady = steadystate(model, usteady)
    zsteady = np.zeros(model.nn_total)
    x0, a, b, c = model.x0.copy(), model.a.copy(), model.b.copy(), model.c.copy()
    y0, dy, ey, fy = model.y0.copy(), model.dy.copy(), model.ey.copy(), model.fy.copy()
    zranges, dzdps, dqlins, eqlins = [], [], [], []
    zoff = 0
    for idx, s in enumerate(model.subs):
        psteady = s.dq @ xsteady + s.eq @ usteady + s.fqprev @ zsteady
        base = _Simple(s.elems, s.fq, s.q0, s.pexp, s.nn, np.zeros(s.np_), s.init_z)
        zsub = homotopy_solve(base, psteady)
        base.set_origin(psteady, zsub)
        if not base.converged():
            raise ValueError(f"Cannot linearize because no solution found at p={psteady}")
        dzdp = -np.linalg.solve(base.last_J, base.last_Jp)
        zsteady[zoff:zoff + s.nn] = zsub
        zr = slice(zoff, zoff + s.nn)
        fqdzdps = [s.fqprev[:, zranges[n]] @ dzdps[n] for n in range(idx)]
        dqlin = s.dq + sum((f @ d for f, d in zip(fqdzdps, dqlins)), np.zeros_like(s.dq))
        eqlin = s.eq + sum((f @ e for f, e in zip(fqdzdps, eqlins)), np.zeros_like(s.eq))
        zranges.append(zr); dzdps.append(dzdp); dqlins.append(dqlin); eqlins.append(eqlin)
        x0 = x0 + c[:, zr] @ (zsub - dzdp @ psteady)
        a = a + c[:, zr] @ dzdp @ dqlin
        b = b + c[:, zr] @ dzdp @ eqlin
        y0 = y0 + fy[:, zr] @ (zsub - dzdp @ psteady)
        dy = dy + fy[:, zr] @ dzdp @ dqlin
        ey = ey + fy[:, zr] @ dzdp @ eqlin
        zoff += s.nn
    return DiscreteModel.from_matrices(a=a, b=b, c=np.zeros((model.nx, 0)), x0=x0, dy=dy, ey=ey,
                                       fy=np.zeros((model.ny, 0)), y0=y0, subs=[], solver=model.solver)
