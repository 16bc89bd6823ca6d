"""Device-side checks the driver's `-m gpu` run carries along: the device exp against libm's over
the full argument range, and compute-sanitizer (memcheck + racecheck) on one small case per kernel family.  The file
sorts last: a surprise here cannot mask the parity tests under `pytest -x`."""
import ctypes as C
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_device_exp_within_two_ulp_of_libm():
    """elements.cuh acme_exp (branch-free, table driven) against the host's exp: the diode / BJT laws are exp-dominated,
    and a Newton iteration count can hinge on the last bits"""
    from acme_jl_b200._lib import check, lib
    n = 1 << 20
    t = np.arange(n) / n
    x = np.where(np.arange(n) % 4 == 0, -760 + 1520 * t, np.where(np.arange(n) % 4 == 1, -40 + 80 * t,
                 np.where(np.arange(n) % 4 == 2, 700 + 50 * t, -700 - 50 * t)))
    x[:10] = [np.nan, np.inf, -np.inf, 0.0, -0.0, 709.782712893384, 709.7827128933841, -745.1332191019412, -745.1332191019411, 1e-320]
    out = np.zeros(n)
    check(lib().acmeb200_diag_exp(x.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), n))
    with np.errstate(over="ignore", under="ignore"):
        ref = np.exp(x)
    same = (out == ref) | (np.isnan(out) & np.isnan(ref))
    ulps = np.abs(out[~same].view(np.int64) - ref[~same].view(np.int64)) if (~same).any() else np.zeros(1, dtype=np.int64)
    assert ulps.max() <= 2, f"max ulp distance {ulps.max()} at x = {x[~same][ulps.argmax()]}"


@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_compute_sanitizer_clean(tool):
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    env = dict(os.environ, SAN_SMALL="1")
    res = subprocess.run([exe, "--tool", tool, "--error-exitcode", "9", sys.executable, os.path.join(ROOT, "tools", "sanitize_run.py")],
                         env=env, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    out = res.stdout + res.stderr
    clean = "ERROR SUMMARY: 0 errors" in out if tool == "memcheck" else "RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings)" in out
    assert res.returncode == 0 and clean, out[-3000:]
