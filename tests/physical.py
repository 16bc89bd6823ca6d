"""Independent cross-check for the waveforms the reference's own tests do NOT pin (SURVEY.md section 8c:
sallenkey / birdie / superover say ``# TODO: further validate y`` in runtests.jl).  TEST INFRASTRUCTURE.

``full_system_run`` steps a circuit WITHOUT the DK-method reduction: at every sample it solves the full implicit
system the reduction starts from (/root/reference/src/ACME.jl:264-273, the `lhs`/`rhs` of `model_matrices`)

    mv v + mi i + (mxd/T + mx/2) x' + mq q = u0 + mu u + (mxd/T - mx/2) x     (element equations, trapezoidal rule)
    Tv v = 0,  Ti i = 0                                                        (Kirchhoff, the topology matrices)
    f(q) = 0                                                                   (element laws)

for w = (v, i, x', q) -- 2 nb + nx + nq unknowns, as many equations -- by a damped Newton iteration in float64
with minimum-norm steps (floating quantities make the Jacobian rank deficient; state and outputs do not depend
on them, ACME.jl:281-283, 303-305), and evaluates  y = pv v + pi i + (px/2 + pxd/T) x' + pq q + (px/2 - pxd/T) x
(ACME.jl:302-311).  None of gensolve, the null-space elimination, the non-linearity decomposition, the matrices
a b c dq eq fq ..., the sub-problem split, the extrapolating / caching / homotopy solvers or step! takes part; what
is shared with the path under test are the element stamps and laws (pinned separately by K4-K9) and the topology.

``sallenkey_lfilter`` is closed form: the textbook unity-gain Sallen-Key low-pass through the bilinear transform.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from acme_jl_b200.hostsolve import eval_table  # noqa: E402


def _f(a):
    return np.array(a, dtype=float)


def full_system_run(circ, fs, u, newton_tol=1e-13, maxiter=400):
    """u: (nu, N) -> y: (ny, N), x: (nx,) final state; raises if a step does not converge."""
    T = 1.0 / fs
    nb, nx, nq, nu, ny = circ.nb, circ.nx, circ.nq, circ.nu, circ.ny
    mv, mi, mx, mxd, mq, mu = (_f(circ.blockdiag(k)) for k in ("mv", "mi", "mx", "mxd", "mq", "mu"))
    pv, pi_, px, pxd, pq = (_f(circ.blockdiag(k)) for k in ("pv", "pi", "px", "pxd", "pq"))
    u0 = _f(circ.u0()).reshape(-1)
    tv, ti = (_f(m).reshape(-1, nb) for m in circ.topomat())
    nl = mv.shape[0]
    table = circ.nl_table()
    nn = sum(e.nn for e, _ in table)
    nw = 2 * nb + nx + nq
    assert nl + nb + nn == nw
    L = np.zeros((nl + nb, nw))
    L[:nl, :nb] = mv
    L[:nl, nb:2 * nb] = mi
    L[:nl, 2 * nb:2 * nb + nx] = mxd / T + mx / 2
    L[:nl, 2 * nb + nx:] = mq
    L[nl:nl + len(tv), :nb] = tv
    L[nl + len(tv):, nb:2 * nb] = ti
    Rx = mxd / T - mx / 2
    Pw = np.hstack([pv, pi_, px / 2 + pxd / T, pq])
    Px = px / 2 - pxd / T
    # row scaling of the linear part (resistances of 1e6 next to unit coefficients)
    scale = 1.0 / np.maximum(np.abs(L).max(axis=1), 1e-300)

    def F(w, rhs):
        res, Jq = eval_table(table, w[2 * nb + nx:], nn)
        return np.concatenate([(L @ w - rhs) * np.concatenate([scale[:nl], np.ones(nb)]), res]), Jq

    x = np.zeros(nx)
    w = np.zeros(nw)
    N = u.shape[1]
    y = np.zeros((ny, N))
    J = np.zeros((nw, nw))
    J[:nl + nb] = L * np.concatenate([scale[:nl], np.ones(nb)])[:, None]
    def newton(w, rhs):
        r, Jq = F(w, rhs)
        for it in range(maxiter):
            J[nl + nb:, :] = 0.0
            J[nl + nb:, 2 * nb + nx:] = Jq
            # two-sided equilibration (amperes next to volts, 1e-12 A/V diode slopes next to 1e6 ohm stamps), then the
            # minimum-norm step: exactly floating quantities give singular values at rounding level, which are cut
            dr = 1.0 / np.maximum(np.abs(J).max(axis=1), 1e-300)
            dc = 1.0 / np.maximum(np.abs(J * dr[:, None]).max(axis=0), 1e-300)
            dw = dc * np.linalg.lstsq(J * dr[:, None] * dc[None, :], -r * dr, rcond=1e-15)[0]
            if np.abs(dw[2 * nb:]).max() <= newton_tol * (1.0 + np.abs(w[2 * nb:]).max()):
                return w + dw, True   # the Newton step itself is at rounding level: converged
            lam, r0 = 1.0, np.abs(r).max()
            while True:  # backtracking on max|F|; exp() overflow counts as "worse"
                with np.errstate(over="ignore", invalid="ignore"):
                    r1, Jq1 = F(w + lam * dw, rhs)
                if np.all(np.isfinite(r1)) and (np.abs(r1).max() <= r0 or lam < 1e-6):
                    break
                lam /= 2
            if lam < 1e-6:
                return w, False
            w = w + lam * dw
            r, Jq = r1, Jq1
            if r0 < 1e-12 and np.abs(r1).max() > 0.5 * r0 and \
                    np.abs(dw[2 * nb:]).max() <= 1e-10 * (1.0 + np.abs(w[2 * nb:]).max()):
                return w, True   # residual at its rounding floor and the steps down to noise in a badly conditioned
                                 # direction (a reverse-biased diode's 1e-12 A/V); larger steps are still refinement
        return w, False

    rhs_prev = np.zeros(nl + nb)   # w = 0 solves the source-free circuit (every element law has f(0) = 0)
    for n in range(N):
        rhs = np.concatenate([u0 + mu @ u[:, n] + Rx @ x, np.zeros(nb)])
        w1, ok = newton(w, rhs)
        if not ok:
            # source stepping from the last solved right-hand side (the supply switching on at sample 0)
            a, da, w1 = 0.0, 0.25, w
            while a < 1.0:
                a_try = min(1.0, a + da)
                w2, ok = newton(w1, rhs_prev + a_try * (rhs - rhs_prev))
                if ok:
                    a, w1, da = a_try, w2, min(2 * da, 0.25)
                else:
                    da /= 2
                    if da < 1e-6:
                        raise RuntimeError(f"full-system Newton did not converge at sample {n}")
        w, rhs_prev = w1, rhs
        y[:, n] = Pw @ w + Px @ x
        x = w[2 * nb:2 * nb + nx].copy()
    return y, x


def dc_solve(circ, u):
    """steady state of the continuous-time circuit for a constant input u: the element equations with x' = x and
    xdot = 0 (mv v + mi i + mx x + mq q = u0 + mu u), Kirchhoff, element laws -- source stepping + damped Newton as
    above.  Returns the state vector x (what steadystate(model, u) computes through (I - a)^-1, ACME.jl:474-497)."""
    nb, nx, nq = circ.nb, circ.nx, circ.nq
    mv, mi, mx, mq, mu = (_f(circ.blockdiag(k)) for k in ("mv", "mi", "mx", "mq", "mu"))
    u0 = _f(circ.u0()).reshape(-1)
    tv, ti = (_f(m).reshape(-1, nb) for m in circ.topomat())
    nl = mv.shape[0]
    table = circ.nl_table()
    nn = sum(e.nn for e, _ in table)
    nw = 2 * nb + nx + nq
    L = np.zeros((nl + nb, nw))
    L[:nl, :nb], L[:nl, nb:2 * nb], L[:nl, 2 * nb:2 * nb + nx], L[:nl, 2 * nb + nx:] = mv, mi, mx, mq
    L[nl:nl + len(tv), :nb] = tv
    L[nl + len(tv):, nb:2 * nb] = ti
    rhs_full = np.concatenate([u0 + mu @ np.asarray(u, float).reshape(-1), np.zeros(nb)])
    J = np.zeros((nw, nw))
    J[:nl + nb] = L

    def newton(w, rhs):
        for it in range(300):
            res, Jq = eval_table(table, w[2 * nb + nx:], nn)
            r = np.concatenate([L @ w - rhs, res])
            J[nl + nb:, :] = 0.0
            J[nl + nb:, 2 * nb + nx:] = Jq
            dr = 1.0 / np.maximum(np.abs(J).max(axis=1), 1e-300)
            dc = 1.0 / np.maximum(np.abs(J * dr[:, None]).max(axis=0), 1e-300)
            dw = dc * np.linalg.lstsq(J * dr[:, None] * dc[None, :], -r * dr, rcond=1e-15)[0]
            if np.abs(dw).max() <= 1e-13 * (1.0 + np.abs(w).max()):
                return w + dw, True
            lam, r0 = 1.0, np.abs(r).max()
            while lam >= 1e-6:
                with np.errstate(over="ignore", invalid="ignore"):
                    w1 = w + lam * dw
                    res1, _ = eval_table(table, w1[2 * nb + nx:], nn)
                    r1 = np.concatenate([L @ w1 - rhs, res1])
                if np.all(np.isfinite(r1)) and np.abs(r1).max() <= r0:
                    break
                lam /= 2
            if lam < 1e-6:
                return w, np.abs(r).max() < 1e-12
            w = w1
        return w, False

    w, a, da = np.zeros(nw), 0.0, 0.25
    while a < 1.0:
        w2, ok = newton(w, min(1.0, a + da) * rhs_full)
        if ok:
            a, w, da = min(1.0, a + da), w2, min(2 * da, 0.25)
        else:
            da /= 2
            if da < 1e-6:
                raise RuntimeError("dc_solve: no convergence")
    return w[2 * nb:2 * nb + nx]


def sallenkey_lfilter(u, fs, r1=10e3, r2=10e3, c1=10e-9, c2=10e-9):
    """unity-gain Sallen-Key low-pass H(s) = 1 / (1 + s c2 (r1 + r2) + s^2 r1 r2 c1 c2) (c1 = feedback capacitor,
    examples/sallenkey.jl:6-17), bilinear transform s = 2 fs (z-1)/(z+1) -- what trapezoidal capacitors amount to"""
    from scipy.signal import bilinear, lfilter
    b, a = bilinear([1.0], [r1 * r2 * c1 * c2, c2 * (r1 + r2), 1.0], fs)
    return lfilter(b, a, u)
