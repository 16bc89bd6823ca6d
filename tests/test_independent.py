"""Waveform parity for the circuits the reference's own tests leave unpinned (sallenkey,
birdie, superover carry `# TODO: further validate y` in runtests.jl), against an INDEPENDENT solve of the circuit
equations (tests/physical.py): a damped full-system Newton on (v, i, x', q) per sample -- no DK reduction, no model
matrices, no extrapolating/caching/homotopy solver -- and, for the linear Sallen-Key filter, the closed-form transfer
function through scipy's bilinear transform.  The oracle (and through it every CUDA parity test) is compared with
these at the north_star's criterion (1e-6 relative, magnitude floored at 1e-3 of the peak); the observed agreement is
1e-9 .. 1e-15 and is asserted at 1e-7.  Newton tolerances are tightened on both sides (set_resabstol!,
solvers.jl:181), as in the birdie parity tests, so that the comparison is not limited by the stopping rule."""
import numpy as np
import pytest

from acme_jl_b200 import BatchRunner, examples as ex
from oracle.oracle import OracleModel

import cases
import physical

H = "HomotopySolver{SimpleSolver}"
HC = "HomotopySolver{CachingSolver{SimpleSolver}}"


def rel_err(y, yref):
    peak = np.abs(yref).max()
    return float(np.max(np.abs(y - yref) / np.maximum(np.abs(yref), 1e-3 * peak)))


def oracle_run(model, u, solver=H):
    return OracleModel(model, 1, solver=solver, tol=1e-13).run(u, threads=0)[:, :, 0]


def sine(n, fs=44100, amp=1.0):
    return (amp * np.sin(2 * np.pi * 1000 / fs * np.arange(n))).reshape(1, -1)


def test_diodeclipper_against_full_system_solve_and_swept_parameters():
    u = sine(400)
    for kw in ({}, dict(is1=1e-13, is2=1.8e-13, η1=1.7, η2=1.7), dict(r=2.2e3, c=22e-9)):
        yp, _ = physical.full_system_run(ex.diodeclipper_circuit(**kw), 44100, u)
        assert rel_err(oracle_run(ex.diodeclipper(**kw), u), yp) < 1e-7


def test_sallenkey_against_transfer_function_and_full_system_solve():
    for fs, kw in ((96000, {}), (44100, {}), (96000, dict(r1=3.3e3, r2=3.3e3, c1=47e-9, c2=4.7e-9))):
        u = sine(600, fs)
        yo = oracle_run(ex.sallenkey(fs=fs, **kw), u)
        assert rel_err(yo[0], physical.sallenkey_lfilter(u[0], fs, **kw)) < 1e-10
        yp, _ = physical.full_system_run(ex.sallenkey_circuit(**kw), fs, u[:, :200])
        assert rel_err(yo[:, :200], yp) < 1e-10


def test_birdie_against_full_system_solve():
    rng = np.random.default_rng(5)
    noise = np.clip(0.2 * rng.standard_normal((1, 400)), -1, 1)          # config 5's input class
    for u in (sine(300, amp=0.3), noise):
        yp, xp = physical.full_system_run(ex.birdie_circuit(0.8), 44100, u)
        for solver in (H, HC):
            assert rel_err(oracle_run(ex.birdie(vol=0.8), u, solver), yp) < 1e-7
    # volume potentiometer as a second input (nu = 2, the [bjt, pot] shape)
    u2 = np.vstack([sine(200, amp=0.3), np.full((1, 200), 0.6)])
    yp, _ = physical.full_system_run(ex.birdie_circuit(), 44100, u2)
    assert rel_err(oracle_run(ex.birdie(), u2), yp) < 1e-7


def test_superover_against_full_system_solve():
    """config 4's circuit, potentiometers as inputs (nn = 13) and baked in (nn = 7): the 9 V supply switching on at
    sample 0 plus a 1 kHz tone"""
    N = 150
    u4 = np.zeros((4, N)); u4[0] = sine(N)[0]; u4[1] = 0.6; u4[2] = 0.4; u4[3] = 1.0
    yp, _ = physical.full_system_run(ex.superover_circuit(), 44100, u4)
    assert np.abs(yp).max() > 0.05                                       # the comparison is not about silence
    assert rel_err(oracle_run(ex.superover(), u4), yp) < 1e-7
    yb, _ = physical.full_system_run(ex.superover_circuit(0.6, 0.4, 1.0), 44100, u4[:1])
    assert rel_err(oracle_run(ex.superover(0.6, 0.4, 1.0), u4[:1]), yb) < 1e-7
    assert rel_err(yb, yp) < 1e-7                                        # both readings of config 4 are one circuit


@pytest.mark.gpu
def test_cuda_path_against_full_system_solve():
    """the CUDA kernels themselves (not via the oracle) against the independent solve: the four BASELINE circuits"""
    N = 120
    u4 = np.zeros((4, N)); u4[0] = sine(N)[0]; u4[1] = 0.6; u4[2] = 0.4; u4[3] = 1.0
    rng = np.random.default_rng(7)
    jobs = [(ex.diodeclipper_circuit(), ex.diodeclipper(), 44100, sine(N)),
            (ex.sallenkey_circuit(), ex.sallenkey(fs=96000), 96000, sine(N, 96000)),
            (ex.birdie_circuit(0.8), ex.birdie(vol=0.8), 44100, np.clip(0.2 * rng.standard_normal((1, N)), -1, 1)),
            (ex.superover_circuit(), ex.superover(), 44100, u4)]
    for circ, model, fs, u in jobs:
        yp, _ = physical.full_system_run(circ, fs, u)
        r = BatchRunner(model, 2, solver=HC, tol=1e-13)
        y = r.run(u)
        r.close()
        assert np.array_equal(y[:, :, 0], y[:, :, 1])
        assert rel_err(y[:, :, 0], yp) < 1e-7, r.kernel_name
