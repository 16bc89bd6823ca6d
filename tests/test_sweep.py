"""Sweep-scale model derivation (acme.jl_b200/sweep.py): the parallel
derivation must equal a plain loop of DiscreteModel(...) calls, and a sweep of baked-in element values
(BASELINE config 3; the alternative reading of config 4) must run through the same ABI."""
import numpy as np
import pytest

import acme_jl_b200 as A
from acme_jl_b200 import examples as ex

import cases


def build_sk(R, kap):
    return ex.sallenkey(fs=96000, r1=R, r2=R, c1=10e-9 * kap, c2=10e-9 / kap)


def build_so(drive, tone):
    return ex.superover(drive, tone, 1.0)


def test_derive_sweep_equals_loop():
    pts = [(10 ** (3 + 2 * k / 7), 10 ** (1.3 * (k % 3) / 2)) for k in range(8)]
    base, kw, B = A.derive_sweep(build_sk, pts, workers=2, chunk=3)
    assert B == 8 and set(kw) == {"overrides"}          # a linear model: nothing but matrices
    ov = kw["overrides"]
    assert set(ov) == {"a", "b", "dy", "ey"}             # x0, y0 do not depend on R, C here
    for b, p in enumerate(pts):
        m = build_sk(*p)
        for key in ("a", "b", "dy", "ey"):
            assert np.array_equal(ov[key][..., b], getattr(m, key))
    serial, kw1, _ = A.derive_sweep(build_sk, pts, workers=1)
    assert all(np.array_equal(kw1["overrides"][k], ov[k]) for k in ov)


def test_derive_sweep_rejects_structure_change():
    # r -> 0 Ohm decouples the two diodes of the clipper: the non-linear problem decomposes into two
    # sub-problems (ACME.jl:277-315); a different model structure in one batch is an error
    def build(r):
        return A.DiscreteModel(ex.diodeclipper_circuit(r=r), 1 / 44100)
    with pytest.raises(ValueError, match="different model structure"):
        A.derive_sweep(build, [1e3, 0.0], workers=1)


def test_derive_sweep_empty():
    with pytest.raises(ValueError, match="empty sweep"):
        A.derive_sweep(build_sk, [])


@pytest.mark.gpu
def test_config3_sweep_through_derive_sweep():
    from acme_jl_b200 import BatchRunner
    from oracle.oracle import OracleModel
    pts = [(10 ** (3 + 2 * (k % 4) / 3), 10 ** (1.3 * (k // 4) / 3)) for k in range(16)]
    base, kw, B = A.derive_sweep(build_sk, pts, workers=1)   # no fork in a process that holds a CUDA context
    u = cases.sine(3000) if hasattr(cases, "sine") else None
    r = BatchRunner(base, B, **kw)
    y = r.run(u)
    yref = OracleModel(base, B, **kw).run(u, threads=0)
    peak = np.abs(yref).max()
    assert np.max(np.abs(y - yref) / np.maximum(np.abs(yref), 1e-3 * peak)) < 1e-6
    assert r.kernel_name.startswith("tpi<linear")
    r.close()


@pytest.mark.gpu
def test_config4_alternative_reading_baked_pots():
    """superover(drive, tone, level) with the potentiometers baked into the stamps (np 5, nn 7,
    runtests.jl:744): per-instance matrices AND per-instance initial solutions, derived at sweep
    scale, against the oracle fed with the same per-instance arrays"""
    from acme_jl_b200 import BatchRunner
    from oracle.oracle import OracleModel
    pts = [(0.15 + 0.2 * (k % 3), 0.3 + 0.4 * (k // 3)) for k in range(6)]
    base, kw, B = A.derive_sweep(build_so, pts, workers=1)
    assert base.subs[0].np_ == 5 and "init_z" in kw and "fq0" in kw["overrides"]
    u = cases.sine(1200)
    H = "HomotopySolver{SimpleSolver}"
    r = BatchRunner(base, B, solver=H, **kw)
    assert r.kernel_name.startswith("rows<") and "per-instance matrices" in r.kernel_name
    y = r.run(u)
    yref = OracleModel(base, B, solver=H, **kw).run(u, threads=0)
    rg = BatchRunner(base, B, solver=H, kernel="generic", **kw)
    yg = rg.run(u)                               # the generic kernel sums J = Jq*fq in another order
    assert np.max(np.abs(yg - y)) <= 1e-9 * np.abs(yref).max()
    rg.close()
    peak = np.abs(yref).max()
    err = np.abs(y - yref) / np.maximum(np.abs(yref), 1e-3 * peak)
    assert err.max() < 1e-6, err.max()
    # and every instance equals its own single-instance model
    for b in (0, 5):
        y1 = BatchRunner(build_so(*pts[b]), 1, solver=H).run(u)
        assert np.max(np.abs(y1[:, :, 0] - y[:, :, b])) <= 1e-9 * peak
    r.close()


@pytest.mark.gpu
def test_steadystate_of_a_baked_sweep():
    """batched steadystate! with per-instance matrices of a non-linear model, on the device"""
    from acme_jl_b200 import BatchRunner
    pts = [(0.2, 0.5), (0.7, 0.4), (0.45, 0.9)]
    base, kw, B = A.derive_sweep(build_so, pts, workers=1)
    r = BatchRunner(base, B, **kw)
    udc = np.array([[0.0, 0.05, -0.02]])
    xs = r.steadystate_(udc)
    for b, p in enumerate(pts):
        assert np.allclose(xs[:, b], build_so(*p).steadystate([udc[0, b]]), rtol=1e-8, atol=1e-11)
    r.run(np.asfortranarray(np.repeat(udc[:, None, :], 64, axis=1)))
    assert np.abs(r.x - xs).max() < 1e-9
    r.close()


@pytest.mark.gpu
def test_linearize_of_a_baked_sweep():
    """batched linearize: every instance's small-signal model == linearize of its own circuit, and the
    linear batch follows the non-linear batch for a small signal around the steady state"""
    from acme_jl_b200 import BatchRunner
    pts = [(0.2, 0.5), (0.7, 0.4), (0.45, 0.9)]
    base, kw, B = A.derive_sweep(build_so, pts, workers=1)
    r = BatchRunner(base, B, **kw)
    lr = r.linearize()
    for b, p in enumerate(pts):
        ref = build_so(*p).linearize()
        for name in ("a", "b", "dy", "ey", "x0", "y0"):
            got = lr._stack(name, getattr(lr.model, name))
            got = got[b if got.shape[0] > 1 else 0]
            want = np.asarray(getattr(ref, name), dtype=float).reshape(got.shape)
            assert np.abs(got - want).max() <= 1e-4 * max(1.0, np.abs(want).max()), name
    N = 512
    u = np.asfortranarray(np.repeat((1e-4 * np.sin(2 * np.pi * 1000 / 44100 * np.arange(N)))[None, :, None], B, axis=2))
    lr.x = lr.steadystate()
    r.steadystate_()
    assert np.abs(lr.run(u) - r.run(u)).max() < 1e-4          # runtests.jl:749's bound for this circuit
    lr.close(); r.close()


@pytest.mark.gpu
def test_batched_element_jacobians_match_the_host_laws():
    """acmeb200_eval_jq: Jq of the whole batch on the device == the host restatement of the element laws
    (circuit.jl:10-17), per instance with its own swept parameters (diodes) and for a BJT stage"""
    from acme_jl_b200 import BatchRunner, hostsolve
    from dataclasses import replace
    rng = np.random.default_rng(3)
    B = 97
    m = ex.diodeclipper()
    P = np.vstack([10.0 ** rng.uniform(-16, -12, B), rng.uniform(1, 2, B), 10.0 ** rng.uniform(-16, -12, B), rng.uniform(1, 2, B)])
    r = BatchRunner(m, B, params=[P])
    s = m.subs[0]
    q = rng.uniform(-0.6, 0.6, (B, s.nq))
    J = r.eval_jq(0, q)
    assert J.shape == (B, s.nn, s.nq)
    for k in range(B):
        table, o = [], 0
        for e, off in s.elems:
            table.append((replace(e, params=tuple(P[o:o + len(e.params), k])), off)); o += len(e.params)
        want = hostsolve.eval_table(table, q[k], s.nn)[1]
        assert np.allclose(J[k], want, rtol=1e-12, atol=1e-300), k
    r.close()
    m = ex.birdie(vol=0.8)
    r = BatchRunner(m, 5)
    s = m.subs[0]
    q = rng.uniform(-0.5, 0.5, (5, s.nq))
    J = r.eval_jq(0, q)
    for k in range(5):
        assert np.allclose(J[k], hostsolve.eval_table(s.elems, q[k], s.nn)[1], rtol=1e-12, atol=1e-300)
    r.close()
