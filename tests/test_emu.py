"""The device library's OWN kernel sources, executed on the CPU.

tests/emu/ holds a host emulation of the slice of CUDA the kernels use (every CUDA thread a fiber,
full-mask warp collectives, host memory for device memory).  tests/emu/build_emu.py compiles
csrc/acmeb200.cu (ABI + generic kernel), csrc/tpi.cu (thread-per-instance kernels, TMA tiles), csrc/coop.cu
(lanes-per-instance kernel) and csrc/rows.cu (the warp-per-instance kernel with the LU rows in registers)
against it with g++; these tests drive that library through the normal Python host layer
(ACMEB200_LIB) in a subprocess and compare with the oracle -- so the warp-level algorithm of the CUDA
kernel (row relabelling instead of swapping, bit-pattern pivot search, augmented right-hand side, the
solver state machine, the rows <-> generic state conversion, per-instance matrices, failure paths) is
checked here, without a GPU.  What the emulation cannot tell -- divergence, bank conflicts, TMA, speed --
is what the `-m gpu` tests and profiles/ are for.  The product never loads this library."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, HC = "HomotopySolver{SimpleSolver}", "HomotopySolver{CachingSolver{SimpleSolver}}"


@pytest.fixture(scope="module")
def emu_lib():
    sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
    import build_emu
    return build_emu.build()


def run_case(lib, case):
    env = dict(os.environ, ACMEB200_LIB=lib)
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "run_case.py"), case], env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    return json.loads(res.stdout.strip().splitlines()[-1])


def test_emulated_rows_kernel_matches_oracle(emu_lib):
    out = run_case(emu_lib, "superover_rows")
    h = out[H]
    assert h["kernel"].startswith("rows<")
    assert h["err"] < 1e-12                                   # same algorithm, same arithmetic up to summation order
    assert h["hist"] == h["hist_ref"] and h["hom"] == h["hom_ref"]   # iteration for iteration, homotopy for homotopy
    assert h["chunked_equal"] and h["generic_diff"] < 1e-12
    c = out[HC]
    assert "dynamic solution cache" in c["kernel"] and c["err"] < 1e-6 and c["chunked_equal"]
    assert sum(abs(a - b) for a, b in zip(c["hist"], c["hist_ref"])) <= 4 and c["hom"] == c["hom_ref"]


def test_emulated_rows_kernel_per_instance_matrices(emu_lib):
    out = run_case(emu_lib, "baked_perinst")
    assert "per-instance matrices" in out["kernel"] and out["err"] < 1e-12 and out["hist"] == out["hist_ref"]


def test_emulated_rows_kernel_four_warp_build(emu_lib):
    out = run_case(emu_lib, "rows_4warp")
    assert "per-instance matrices" in out["perinst"]["kernel"] and out["perinst"]["err"] < 1e-6 and out["perinst"]["samples"] == 6 * 40
    assert out["shared"]["kernel"].startswith("rows<") and out["shared"]["err"] < 1e-12 and out["shared"]["samples"] == 5 * 30


def test_batched_steadystate_with_per_instance_matrices(emu_lib):
    """steadystate / steadystate! for a sweep of baked-in element values (host logic + the ABI, here against the
    emulated library): equals each instance's own host-side steadystate, and the instances stay there"""
    out = run_case(emu_lib, "steady")
    assert out["shared"] < 1e-10 and out["perinst"] < 1e-10 and out["linear"] < 1e-10
    assert out["reuse"] < 1e-10          # a second call on the same runner reuses (and resets) the derived device model
    assert out["drift"] < 1e-9          # checksteady! (runtests.jl:664-682) compares the state
    assert out["y_span"] < 1e-4          # the output only up to the Newton stopping rule (res < 1e-10 A at high-impedance nodes)


def test_batched_linearize(emu_lib):
    """BatchRunner.linearize (ACME.jl:505-550 per instance): per-instance steady inputs, element parameters and
    baked matrices give each instance the small-signal model of its own separately built circuit (up to the
    host reference's 1e-10 steady-state tolerance; the device solves to 1e-15), and the linear batch tracks
    the non-linear one for a small signal (runtests.jl:673-682)"""
    out = run_case(emu_lib, "linearize")
    assert out["inputs"] < 1e-9 and out["params"] < 1e-7 and out["baked"] < 1e-4
    assert out["tracking"] < 1e-7
    assert out["shared_has_overrides"] is False


def test_emulated_rows_kernel_failure_semantics(emu_lib):
    out = run_case(emu_lib, "failure")
    for solver in (H, "SimpleSolver"):
        o = out[solver]
        assert o["status"] == o["status_ref"] and o["first"] == o["first_ref"] and o["hist"] == o["hist_ref"]
    assert any(out["SimpleSolver"]["status"])                 # the plain solver does fail on this drive


def test_emulated_tpi_kernel_golden_vector_and_batches(emu_lib):
    """k_tpi -- the headline kernel -- under emulation: 2D tensor-map tiles with their swizzle, mbarrier phases,
    hot/cold step split, learning cache.  G1 is the reference's own doctest output."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    out = run_case(emu_lib, "tpi")
    g = cases.golden()["G1_diodeclipper_doctest"]
    assert out["g1"]["kernel"].startswith("tpi<diodeclipper") and out["g1"]["n"] == g["n"]
    for got, want in zip(out["g1"]["first"], g["first"]):
        cases.assert_printed_equal(got, want, g["printed_digits"])
    for got, want in zip(out["g1"]["last"], g["last"]):
        cases.assert_printed_equal(got, want, g["printed_digits"])
    assert out["clipper"]["err"] < 1e-12 and out["clipper"]["hist"] == out["clipper"]["hist_ref"]
    assert out["chunks"]["equal"] and out["chunks"]["launches"] >= 4   # init + 3 time chunks, bit-identical to one chunk
    assert out["linear"]["kernel"].startswith("tpi<linear") and out["linear"]["err"] < 1e-13
    assert out["birdie"]["kernel"].startswith("tpi<birdie") and out["birdie"]["err"] < 1e-6
    assert out["birdie"]["bad"] == 0 and max(out["birdie"]["stored"]) > 1


def test_emulated_sample_major_streams(emu_lib):
    """ACMEB200_SAMPLE_MAJOR under emulation: transposed tensor-map tiles (box {32*channels, T} of the (channels*B, N)
    stream), the synchronous path for odd pitches, the host time-chunk pipeline, linear whole-tile path, two input
    channels, mixed-layout calls -- each bit-identical to the default layout, which the tests above pin to the oracle"""
    out = run_case(emu_lib, "sample_major")
    for key, samples in (("clipper", 70 * 203), ("odd_batch", 37 * 203), ("chunks", 70 * 5003), ("linear", 38 * 101), ("birdie_vol", 6 * 150)):
        o = out[key]
        assert o["kernel"].startswith("tpi<") and o["equal"] and o["hist_equal"] and o["samples"] == samples, (key, o)
    assert out["chunks"]["launches"] >= 4                     # init + 3 time chunks
    assert out["clipper"]["err"] < 1e-6 and out["linear"]["err"] < 1e-13
    assert out["shared_u"] and out["mixed_calls"] and out["rows_refused"]
    g = out["generic"]
    assert g["kernel"].startswith("generic<") and g["equal"] and g["hist_equal"] and g["samples"] == 37 * 203


def test_emulated_cooperative_kernel(emu_lib):
    """k_coop under emulation, including its sub-warp masks (two 16-lane groups per warp with independent control flow)"""
    out = run_case(emu_lib, "coop")
    st = out["static"]
    assert st["kernel"].startswith("coop<32") and st["err"] < 1e-6 and st["equals_rows"] and st["hist_equals_rows"]
    ms = out["multisub"]
    assert ms["kernel"].startswith("coop<16") and "runtime dims" in ms["kernel"] and ms["nsub"] == 3
    assert ms["err"] < 1e-12 and ms["all_instances_equal"] and ms["hist"] == ms["hist_ref"]


def test_gpu_parity_suite_subset_under_emulation(emu_lib):
    """The `-m gpu` parity tests themselves, driven against the emulated library: the known-answer tests of
    runtests.jl (K3-K9), frozen k-d tree caches, batched steady state, unaligned/odd streams (the synchronous
    tile path), per-instance matrices of a non-linear model, size checks.  (The long-running ones -- full
    waveforms on the lane-parallel kernels -- stay GPU-only.)"""
    sel = ("K4 or K5 or K6 or K7 or K8 or K9 or empty_circuits or io_size or frozen_cache_lookup or steadystate_on_device or "
           "run_bang or odd_lengths or per_instance_matrices_nonlinear or (K3 and not coop) or K12 or K13 or (solver_state and generic)")
    env = dict(os.environ, ACMEB200_LIB=emu_lib)
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x",
                          "-p", "no:cacheprovider", "-k", sel], env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = res.stdout.strip().splitlines()[-1] if res.stdout.strip() else res.stderr[-500:]
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-1000:]
    assert " passed" in tail and int(tail.split(" passed")[0].split()[-1]) >= 20, tail


def test_sample_major_gpu_tests_under_emulation(emu_lib):
    """tests/test_z_sample_major.py (all but the torch-device-tensor one) against the emulated library"""
    env = dict(os.environ, ACMEB200_LIB=emu_lib)
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_z_sample_major.py"), "-m", "gpu", "-q", "-x",
                          "-p", "no:cacheprovider", "-k", "not device_tensors"], env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0 and "6 passed" in res.stdout, res.stdout[-3000:] + res.stderr[-1000:]


def test_cuda_kernels_against_independent_full_system_solve_under_emulation(emu_lib):
    """tests/test_independent.py's GPU test -- the four BASELINE circuits through the kernels' own sources against a
    full-system Newton solve of the circuit equations (no DK reduction, no oracle) -- on the emulated library"""
    env = dict(os.environ, ACMEB200_LIB=emu_lib)
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_independent.py"), "-m", "gpu", "-q", "-x",
                          "-p", "no:cacheprovider"], env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0 and "1 passed" in res.stdout, res.stdout[-3000:] + res.stderr[-1000:]


def test_batched_element_jacobians(emu_lib):
    """acmeb200_eval_jq (the element Jacobians of a whole batch in one launch, what the batched linearize uses)
    against the host restatement of the element laws: swept diode parameters, and the five elements of superover"""
    out = run_case(emu_lib, "eval_jq")
    assert out["clipper"] < 1e-12 and out["superover"] < 1e-9 and out["shape"]

