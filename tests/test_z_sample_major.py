"""Sample-major streams (ACMEB200_SAMPLE_MAJOR; SURVEY.md section 8(d) "instance-fastest" layout) through the C ABI
on the GPU: bit-identical to the default layout and in parity with the oracle.  Kept in a file of its own that sorts
after the established parity tests: these were written after round 1's last GPU visit and have so far run under the
host emulation only (tests/test_emu.py), so with `pytest -x` a surprise here cannot mask the rest."""
import numpy as np
import pytest

import acme_jl_b200 as A
from acme_jl_b200 import BatchRunner, examples as ex
from oracle.oracle import OracleModel

import cases
from test_gpu_parity import H, HC, assert_parity, clipper_sweep

pytestmark = pytest.mark.gpu


def _smaj(a):
    """(channels, N, B) -> (channels, B, N), Fortran order: the sample-major stream of the same data"""
    return np.asfortranarray(np.transpose(a, (0, 2, 1)))


@pytest.mark.parametrize("B,N", [(70, 203), (37, 333), (64, 1), (1, 50)])
def test_sample_major_streams_nonlinear(B, N):
    """ACMEB200_SAMPLE_MAJOR ("instance-fastest" layout): (nu, B, N) / (ny, B, N) host
    streams through the transposed tiles of the thread-per-instance kernel -- bit-identical to the default
    layout and in parity with the oracle.  Odd B: the sample pitch breaks TMA's 16-byte rule -> synchronous path."""
    m = ex.diodeclipper()
    P = clipper_sweep(B)
    u = np.asfortranarray(cases.sine(N)[:, :, None] * np.linspace(0.2, 1.2, B)[None, None, :])
    yref = OracleModel(m, B, params=[P], solver=HC).run(u, threads=0)
    r = BatchRunner(m, B, params=[P], solver=HC)
    y = r.run(u)
    h = r.stats()["iter_hist"]
    r.reset()
    ys = r.run(_smaj(u), layout="sample")
    assert ys.shape == (1, B, N) and np.array_equal(ys, _smaj(y))
    assert r.stats()["iter_hist"] == h
    assert_parity(np.transpose(ys, (0, 2, 1)), yref)
    # one shared input is the same in both layouts; caller-provided output array
    r.reset()
    y1 = r.run(cases.sine(N))
    r.reset()
    yo = np.zeros((1, B, N), order="F")
    assert r.run(cases.sine(N), yo, layout="sample") is yo and np.array_equal(yo, _smaj(y1))
    r.close()


def test_sample_major_streams_linear_and_two_inputs():
    """the whole-tile register path of the linear kernel (per-instance matrices) and 512-byte tile rows (nu = 2)"""
    base, kw, B = A.derive_sweep(lambda R: ex.sallenkey(fs=96000, r1=R, r2=R), [1e3 * (1 + k) for k in range(38)], workers=1)
    u = np.asfortranarray(cases.sine(101)[:, :, None] * (1 + np.arange(B))[None, None, :])
    yref = OracleModel(base, B, **kw).run(u, threads=0)
    r = BatchRunner(base, B, **kw)
    assert r.kernel_name.startswith("tpi<linear")
    y = r.run(u)
    r.reset()
    ys = r.run(_smaj(u), layout="sample")
    r.close()
    assert np.array_equal(ys, _smaj(y))
    assert_parity(y, yref, rtol=1e-12)
    m = ex.birdie()
    Bb, Nb = 6, 150
    rng = np.random.default_rng(3)
    ub = np.zeros((2, Nb, Bb), order="F")
    ub[0] = np.clip(0.2 * rng.standard_normal((Nb, Bb)), -1, 1)
    ub[1] = (0.3 + 0.1 * np.arange(Bb))[None, :]
    r = BatchRunner(m, Bb, solver=HC)
    assert r.kernel_name.startswith("tpi<birdie nx3 nu2")
    y = r.run(ub)
    r.reset()
    ys = r.run(_smaj(ub), layout="sample")
    r.close()
    assert np.array_equal(ys, _smaj(y))


def test_sample_major_state_persists_and_errors():
    """state carries over between calls in either layout; strides are checked; the generic kernel takes the flag,
    the lane-parallel kernels refuse it (no silent relayout)"""
    m = ex.diodeclipper()
    B, N = 40, 300
    P = clipper_sweep(B)
    u = np.asfortranarray(cases.sine(N)[:, :, None] * np.linspace(0.2, 1.2, B)[None, None, :])
    r = BatchRunner(m, B, params=[P], solver=H)
    y = r.run(u)
    r.reset()
    ya = r.run(_smaj(u[:, :100]), layout="sample")
    yb = r.run(np.asfortranarray(u[:, 100:]))
    assert np.array_equal(np.concatenate([np.transpose(ya, (0, 2, 1)), yb], axis=1), y)
    with pytest.raises(A.DimensionMismatch):
        r.run(_smaj(u)[:, :-1], layout="sample")                      # B-1 instances
    with pytest.raises(ValueError):
        r.run(u, layout="rows")
    r.close()
    # the generic thread-per-instance kernel takes the flag (same bits) ...
    rg = BatchRunner(m, B, params=[P], solver=H, kernel="generic")
    yg = rg.run(u)
    rg.reset()
    assert np.array_equal(rg.run(_smaj(u), layout="sample"), _smaj(yg))
    rg.close()
    # ... the lane-parallel kernels (here: rows) refuse it
    us = np.zeros((4, 8, 2), order="F"); us[1:] = 0.5
    rr = BatchRunner(ex.superover(), 2, solver=H)
    with pytest.raises(Exception, match="thread-per-instance"):
        rr.run(_smaj(us), layout="sample")
    rr.close()


def test_sample_major_device_tensors():
    """device pointers (torch (N, B, nu) / (N, B, ny) tensors), asynchronous on the caller's stream; the host-buffer
    time-chunk pipeline gives the same bits"""
    import torch
    B, N = 96, 2048
    m = ex.diodeclipper()
    P = clipper_sweep(B)
    r = BatchRunner(m, B, params=[P], solver=HC)
    u_host = np.asfortranarray(cases.sine(N)[:, :, None] * np.linspace(0.2, 1.2, B)[None, None, :])
    y_host = r.run(u_host)
    r.reset()
    u_dev = torch.from_numpy(np.ascontiguousarray(u_host.transpose(1, 2, 0))).cuda()   # (N, B, nu)
    y_dev = r.run(u_dev, layout="sample")
    torch.cuda.synchronize()
    assert tuple(y_dev.shape) == (N, B, 1)
    assert np.array_equal(y_dev.cpu().numpy().transpose(2, 0, 1), y_host)               # -> (ny, N, B)
    r.reset()
    assert np.array_equal(r.run(_smaj(u_host), layout="sample"), _smaj(y_host))
    r.close()

