#!/usr/bin/env python
"""Headline benchmark: batched circuit-samples per second (BASELINE.json metric).

Workload (BASELINE.json configs[1]): examples/diodeclipper.jl, batch = 65 536
instances per GPU with swept Is / eta (256 x 256 grid, SURVEY.md section 8d
config 2), 1 s of a unit 1 kHz sine at 44.1 kHz per instance.  One "step" = one
``run!`` of that second for the whole batch.  Weak scaling: every GPU owns its
own 65 536 instances (independent instances shard with no data-path collective).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...            # CPU restatement of the reference

Prints ONE JSON line (rank 0).  `value` is timed with the streams resident in
HBM (CUDA events on the launching stream); `e2e` is the same metric through the
public API with pinned HOST buffers, H2D and D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 44100
N_SAMPLES = 44100
BATCH_PER_GPU = 65536
SOLVER = "HomotopySolver{CachingSolver{SimpleSolver}}"  # the reference's default (ACME.jl:150)
METRIC = "Msamples/sec (batched circuit-samples)"
ALG_BYTES_PER_SAMPLE = 16  # one f64 input read + one f64 probe sample written


def sweep_params(batch_total: int, first: int, count: int) -> np.ndarray:
    """Is_k = 10^(-16 + 4k/255), eta_j = 1 + j/255, d2 uses 1.8*Is (diodeclipper.jl:11-12)."""
    idx = np.arange(first, first + count) % 65536
    k, j = idx % 256, idx // 256
    Is = 10.0 ** (-16 + 4 * k / 255)
    eta = 1 + j / 255
    return np.vstack([Is, eta, 1.8 * Is, eta])


def sine_row() -> np.ndarray:
    return np.sin(2 * np.pi * 1000 / FS * np.arange(N_SAMPLES))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.t.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
            except Exception:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cores() -> int:
    """usable host cores: CPU affinity, capped by the cgroup CPU quota if there is one"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(np.ceil(int(quota) / int(period)))))
    except Exception:
        pass
    return n


def cpu_baseline(budget_s: float = 12.0, threads: int = 0) -> dict:
    """The oracle (CPU restatement of the reference's run! path) on all host cores, on a
    bounded sample of the same workload: the first `b` instances of the sweep, full 1 s signal."""
    from acme_jl_b200 import examples as ex
    from oracle import oracle
    from oracle.oracle import OracleModel
    cores = threads or host_cores()
    m = ex.diodeclipper()
    u = sine_row().reshape(1, -1)
    b = 4 * cores
    P = sweep_params(BATCH_PER_GPU, 0, b)
    o = OracleModel(m, b, params=[P], solver=SOLVER)
    t0 = time.perf_counter(); o.run(u, threads=cores); dt = time.perf_counter() - t0
    rate = b * N_SAMPLES / dt
    b2 = int(min(BATCH_PER_GPU, max(b, budget_s * rate / N_SAMPLES)))
    b2 -= b2 % cores or 0
    b2 = max(b2, cores)
    idx = (np.arange(b2) * (BATCH_PER_GPU // b2)) % BATCH_PER_GPU  # spread over the sweep grid
    P = sweep_params(BATCH_PER_GPU, 0, BATCH_PER_GPU)[:, idx]
    o = OracleModel(m, b2, params=[P], solver=SOLVER)
    t0 = time.perf_counter(); o.run(u, threads=cores); dt = time.perf_counter() - t0
    return {"value": b2 * N_SAMPLES / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": f"{b2} of the 65536 swept instances x {N_SAMPLES} samples, {dt:.1f} s, "
                      f"HomotopySolver{{CachingSolver{{SimpleSolver}}}} restated in C (oracle/acme_oracle.c); "
                      f"the Julia reference cannot run here (no Julia toolchain)"}


def measured_peak() -> tuple:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def run_reference(args, rank: int, world: int, emit=print):
    """--impl reference: the reference's CPU algorithm (restated; see cpu_baseline) on the host cores."""
    if rank != 0:
        return
    per_step, per_ms = [], []
    info = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        info = cpu_baseline(budget_s=6.0)
        if i >= args.warmup:
            per_step.append(info["value"])
            per_ms.append((time.perf_counter() - t0) * 1e3)
    v = float(np.mean(per_step)) if per_step else info["value"]
    info["value"] = v
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(per_ms)) if per_ms else None,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "diodeclipper.jl, swept Is/eta, 1 s of 1 kHz sine @ 44.1 kHz; each step is a bounded "
                               "sample (a few seconds of host work) of that sweep, calibration run included in ms_per_step",
                   "batch_per_gpu": BATCH_PER_GPU, "samples": N_SAMPLES},
        "cpu_baseline": info,
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: superover.jl, batch = 8192 with swept drive/tone pots, 1 s @ 44.1 kHz,
# sharded across the GPUs (SURVEY.md section 8d "Config 4"): STRONG scaling, no data-path
# collective; the final NCCL all-gather of the output shards is timed separately.
C4_BATCH = 8192
C4_ALG_BYTES = 40  # API-literal: 4 input rows (audio + 3 constant pot rows) read + 1 probe sample written


def c4_inputs_np(first: int, count: int, n: int) -> np.ndarray:
    """(count, n, 4): unit 1 kHz sine, drive_k = (k+1/2)/128, tone_j = (j+1/2)/64, level = 1"""
    idx = np.arange(first, first + count)
    u = np.empty((count, n, 4))
    u[:, :, 0] = np.sin(2 * np.pi * 1000 / FS * np.arange(n))[None, :]
    u[:, :, 1] = ((idx % 128 + 0.5) / 128)[:, None]
    u[:, :, 2] = ((idx // 128 % 64 + 0.5) / 64)[:, None]
    u[:, :, 3] = 1.0
    return u


def c4_cpu_baseline(budget_s: float = 12.0) -> dict:
    from acme_jl_b200 import examples as ex
    from oracle.oracle import OracleModel
    cores = host_cores()
    m = ex.superover()
    b, n = 2 * cores, 4410
    u = np.asfortranarray(c4_inputs_np(0, C4_BATCH, n)[:: C4_BATCH // b][:b].transpose(2, 1, 0))
    o = OracleModel(m, b, solver=SOLVER)
    t0 = time.perf_counter(); o.run(u, threads=cores); dt = time.perf_counter() - t0
    n2 = int(min(N_SAMPLES, max(n, n * budget_s / max(dt, 1e-3))))
    u = np.asfortranarray(c4_inputs_np(0, C4_BATCH, n2)[:: C4_BATCH // b][:b].transpose(2, 1, 0))
    o = OracleModel(m, b, solver=SOLVER)
    t0 = time.perf_counter(); o.run(u, threads=cores); dt = time.perf_counter() - t0
    return {"value": b * n2 / dt / 1e6, "unit": "Msamples/s", "cores": cores, "kind": "port",
            "sample": f"{b} of the 8192 swept instances x {n2} samples from x = 0, {dt:.1f} s, {SOLVER} restated in C "
                      f"(oracle/acme_oracle.c); the Julia reference cannot run here (no Julia toolchain)"}


# ---------------------------------------------------------------------------------------------
# helpers shared by the per-config measurements
class Dist:
    """rank / world plumbing: barrier, max over ranks, broadcast of small host arrays"""

    def __init__(self, rank, world, dev):
        self.rank, self.world, self.dev = rank, world, dev

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()

    def maxr(self, x: float) -> float:
        if self.world == 1:
            return float(x)
        import torch
        import torch.distributed as dist
        t = torch.tensor([x], dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sumr(self, arr):
        """element-wise sum of a list of numbers over the ranks (whole-job counters)"""
        import torch
        if self.world == 1:
            return [float(a) for a in arr]
        import torch.distributed as dist
        t = torch.tensor(list(arr), dtype=torch.float64, device=self.dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(v) for v in t.tolist()]

    def bcast_arrays(self, arrays):
        """dict of float64 numpy arrays from rank 0 to every rank"""
        if self.world == 1:
            return arrays
        import torch.distributed as dist
        box = [arrays if self.rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]


def timed_steps(D: Dist, stream, fn, warmup: int, steps: int) -> float:
    """`warmup` untimed calls of fn, then `steps` timed ones: CUDA events on the launching stream, barrier +
    synchronize on both sides, max over ranks; returns milliseconds per step"""
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    D.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize(); D.barrier()
    return D.maxr(e0.elapsed_time(e1)) / steps


def host_copy_ceiling(D: Dist) -> dict:
    """What the host side of this box can move: pinned host <-> device copies of 1 GiB, both directions at once on two
    streams, every rank at the same time (the e2e leg does exactly that, plus a kernel): the e2e rate cannot exceed
    aggregate_GBs_per_direction / 8 bytes per sample and direction."""
    import torch
    n = 1 << 27  # doubles: 1 GiB
    hu = torch.empty(n, dtype=torch.float64, pin_memory=True); hy = torch.empty(n, dtype=torch.float64, pin_memory=True)
    du = torch.empty(n, dtype=torch.float64, device=D.dev); dy = torch.zeros(n, dtype=torch.float64, device=D.dev)
    hu.zero_()
    s1, s2 = torch.cuda.Stream(D.dev), torch.cuda.Stream(D.dev)
    reps = 4
    for timed in (False, True):
        torch.cuda.synchronize(); D.barrier()
        t0 = time.perf_counter()
        for _ in range(reps if timed else 1):
            with torch.cuda.stream(s1):
                du.copy_(hu, non_blocking=True)
            with torch.cuda.stream(s2):
                hy.copy_(dy, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    dt = D.maxr(dt)
    per_dir = n * 8 * reps / dt / 1e9
    del hu, hy, du, dy
    return {"GBs_per_direction_per_gpu": per_dir, "aggregate_GBs_per_direction": per_dir * D.world,
            "Msamples_per_s_ceiling": per_dir * D.world / 8 * 1e3,
            "how": "1 GiB pinned H2D and D2H concurrently on two streams, all ranks at once, 4 repetitions"}


def rel_err(y: np.ndarray, yref: np.ndarray) -> float:
    """the parity measure of tests/test_gpu_parity.py: |y - yref| / max(|yref|, 1e-3 max|yref|)"""
    peak = float(np.max(np.abs(yref))) if yref.size else 0.0
    return float(np.max(np.abs(y - yref) / np.maximum(np.abs(yref), max(1e-3 * peak, 1e-300)))) if yref.size else 0.0


def parity_spot(model, B, u_np, solver, what, **kw) -> dict:
    """Outside every timed region: a FRESH device runner and the oracle on the same few instances and inputs, both
    from the initial state.  `u_np` is (nu, N) shared or (nu, N, B) Fortran-ordered.  With the default solver the
    reference's own stopping rule (max|res| < 1e-10, solvers.jl:226) leaves its output uncertain by e_ref (oracle at
    the default tolerance against the oracle at 1e-13), reported beside the error."""
    from acme_jl_b200 import BatchRunner
    from oracle.oracle import OracleModel
    r = BatchRunner(model, B, solver=solver, **kw)
    y = r.run(u_np, check_status=False)
    st = r.stats(); bad = int((r.status()[0] != 0).sum())
    r.close()
    o = OracleModel(model, B, solver=solver, **kw)
    yref = o.run(u_np, threads=0)
    so = o.stats()
    out = {"max_rel_err": rel_err(y, yref), "instances": B, "samples": int(u_np.shape[1]), "against": "oracle (C restatement of the reference), "
           "fresh runners on both sides, outside the timed region", "what": what, "tolerance": 1e-6,
           "newton_iters_device": int(st["newton_iters"]), "newton_iters_oracle": int(so["newton_iters"]), "status_nonzero": bad}
    if model.subs:
        yexact = OracleModel(model, B, solver="HomotopySolver{SimpleSolver}", tol=1e-13, **kw).run(u_np, threads=0)
        out["e_ref_rel"] = rel_err(yref, yexact)
        out["rel_err_vs_converged"] = rel_err(y, yexact)
    # the same two bars as tests/test_gpu_parity.py (assert_parity / assert_parity_within_reference_accuracy)
    e_ref = out.get("e_ref_rel", 0.0)
    if out["max_rel_err"] <= out["tolerance"]:
        out["verdict"] = "strict: within the stated tolerance"
    elif out["max_rel_err"] <= out["tolerance"] + 1.5 * e_ref and out.get("rel_err_vs_converged", 0.0) <= 1.5 * e_ref + out["tolerance"]:
        out["verdict"] = ("within the reference's own stopping uncertainty e_ref_rel (Newton stops at max|res| < 1e-10, solvers.jl:226: two "
                          "faithful implementations may stop one iteration apart; the device result is as close to the converged solution as the reference's)")
    else:
        out["verdict"] = "FAIL"
    return out


def src_stamp() -> str:
    """content hash of the kernel sources: profiles/*.json captures carry the stamp of the build they measured"""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "acme.jl_b200", "csrc")
    for f in sorted(os.listdir(d)):
        h.update(f.encode()); h.update(open(os.path.join(d, f), "rb").read())
    return h.hexdigest()[:16]


def ncu_capture(name: str):
    """an ncu capture summary under profiles/ (written by tools/ncu_refresh.py on the GPU box), or None when it is
    missing or was taken from other kernel sources than the ones in this tree"""
    p = os.path.join(ROOT, "profiles", name)
    try:
        c = json.load(open(p))
    except Exception:
        return None
    if c.get("src_stamp") != src_stamp():
        return None
    return c


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[0]: examples/diodeclipper.jl, ONE instance, 1 s of 1 kHz sine @ 44.1 kHz on the CPU
def config1_record() -> dict:
    """the reference's own CPU-runnable case (runtests.jl:700 / gettingstarted.md doctest) on one host core"""
    from acme_jl_b200 import examples as ex
    from oracle.oracle import OracleModel
    m = ex.diodeclipper()
    u = np.sin(2 * np.pi * 1000 / 44100 * np.arange(44100)).reshape(1, -1)
    best = 1e30
    for _ in range(5):
        o = OracleModel(m, 1, solver=SOLVER)
        t0 = time.perf_counter(); y = o.run(u, threads=1); best = min(best, time.perf_counter() - t0)
    return {"workload": "examples/diodeclipper.jl, 1 instance, 1 s of 1 kHz sine @ 44.1 kHz (BASELINE.json configs[0])",
            "value": 44100 / best / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port", "ms": best * 1e3, "solver": SOLVER,
            "golden_last3": [float(v) for v in y[0, -3:, 0]],
            "note": "oracle (C restatement), best of 5; the doctest's printed samples are -0.537508, -0.462978, -0.36521"}


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[2]: sallenkey.jl low-pass, batch = 65 536 with swept R/C, 1 s @ 96 kHz
C3_BATCH, C3_N, C3_FS = 65536, 96000, 96000


def c3_points(batch: int):
    """R_k = 10^(3 + 2k/255), kappa_j = 10^(1.3 j/255); instance index = 256 j + k"""
    return [(10 ** (3 + 2 * (i % 256) / 255), 10 ** (1.3 * (i // 256 % 256) / 255)) for i in range(batch)]


def c3_derive(batch: int) -> dict:
    """host side (once per model, before CUDA is touched: the derivation pool forks): the exact-rational derivation
    of every swept Sallen-Key model"""
    import acme_jl_b200 as A
    from acme_jl_b200 import examples as ex
    t0 = time.perf_counter()
    _, kw, _ = A.derive_sweep(lambda R, kap: ex.sallenkey(fs=C3_FS, r1=R, r2=R, c1=10e-9 * kap, c2=10e-9 / kap),
                              c3_points(batch), chunk=128)
    return {"overrides": kw["overrides"], "derive_s": time.perf_counter() - t0}


def config3_record(args, D: Dist, pre) -> dict:
    import torch
    from acme_jl_b200 import BatchRunner, examples as ex
    B, N = args.c3_batch, C3_N
    pre = D.bcast_arrays(pre)
    ov = pre["overrides"]
    base = ex.sallenkey(fs=C3_FS)
    runner = BatchRunner(base, B, overrides=ov)
    stream = torch.cuda.current_stream(D.dev)
    row = torch.from_numpy(np.sin(2 * np.pi * 1000 / C3_FS * np.arange(N))).to(D.dev)
    peak, peak_src = measured_peak()
    res = {}
    for layout in ("sample", "instance"):   # 100 GB of streams per layout: one at a time
        if layout == "sample":
            U = row.reshape(N, 1, 1).expand(N, B, 1).contiguous(); Y = torch.empty_like(U)
        else:
            U = row.reshape(1, N, 1).expand(B, N, 1).contiguous(); Y = torch.empty_like(U)
        l0 = runner.launch_count
        ms = timed_steps(D, stream, lambda: runner.run(U, Y, check_status=False, layout=layout), 3, args.sub_steps)
        rate = B * N / (ms / 1e3)          # per GPU
        res[layout] = {"ms_per_step": ms, "value": rate * D.world / 1e6, "achieved_GBs": ALG_BYTES_PER_SAMPLE * rate / 1e9,
                       "frac": ALG_BYTES_PER_SAMPLE * rate / 1e9 / peak, "launches": int(runner.launch_count - l0),
                       "checksum": float(Y[777 % B, N - 1, 0] if layout == "instance" else Y[N - 1, 777 % B, 0])}
        del U, Y
        torch.cuda.empty_cache()
    kernel = runner.kernel_name
    runner.close()
    assert res["sample"]["checksum"] == res["instance"]["checksum"], "the two stream layouts must give identical results"
    out = None
    if D.rank == 0:
        best = res["sample"]
        spots = sorted({i % B for i in (0, 255, 256 * 255, B - 1, 31337, 777)})
        ovs = {k: np.ascontiguousarray(v[..., spots]) for k, v in ov.items()}
        par = parity_spot(base, len(spots), np.sin(2 * np.pi * 1000 / C3_FS * np.arange(N)).reshape(1, -1), SOLVER,
                          f"{len(spots)} instances spread over the (R, kappa) grid, the full second", overrides=ovs)
        cap = ncu_capture("r2/traffic_linear.json")
        out = {"workload": f"examples/sallenkey.jl topology, batch={B} instances/GPU with swept R (256) x kappa (256): per-instance "
                           f"matrices a,b,dy,ey derived on the host in exact rationals, 1 s of unit 1 kHz sine @ 96 kHz "
                           f"(BASELINE.json configs[2]); weak scaling over the GPUs",
               "value": best["value"], "unit": "Msamples/s", "ms_per_step": best["ms_per_step"], "steps": args.sub_steps, "warmup": 3,
               "scaling": "weak", "batch_per_gpu": B, "samples": N, "kernel": kernel, "layout": "sample-major (nu,B,N) streams (ACMEB200_SAMPLE_MAJOR)",
               "l2": "streams larger than L2 (50 GB U + 50 GB Y per GPU per step)",
               "roofline": {"bound": "hbm", "achieved": best["achieved_GBs"], "peak": peak, "unit": "GB/s", "frac": best["frac"],
                            "traffic": cap.get("dram_bytes_per_launch") if cap else None, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": ALG_BYTES_PER_SAMPLE * B * N, "kernel_ms": best["ms_per_step"]},
               "instance_major": {k: res["instance"][k] for k in ("value", "ms_per_step", "achieved_GBs", "frac")},
               "gpu_launches": best["launches"], "host_derivation_s": pre["derive_s"], "parity": par}
    return out


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[3]: superover.jl, batch = 8192 sharded over the GPUs (strong scaling)
def config4_record(args, D: Dist, steps: int, warmup: int, want_e2e: bool, want_cpu: bool) -> dict:
    import torch
    from acme_jl_b200 import examples as ex
    from acme_jl_b200.distributed import ShardedBatchRunner
    dev, rank, world = D.dev, D.rank, D.world
    workload = ("examples/superover.jl (pots as inputs), batch=8192 with swept drive (128) x tone (64), level 1, "
                "1 s of unit 1 kHz sine @ 44.1 kHz per instance, sharded over the GPUs (BASELINE.json configs[3])")
    n = N_SAMPLES
    model = ex.superover()
    sharded = ShardedBatchRunner(model, C4_BATCH, rank=rank, world=world, solver=SOLVER, kernel=args.kernel)
    runner, first, count = sharded.runner, sharded.first, sharded.count
    U = torch.from_numpy(c4_inputs_np(first, count, n)).to(dev)
    Y = torch.empty((count, n, 1), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)
    for _ in range(warmup):
        runner.run(U, Y, check_status=False)
    torch.cuda.synchronize()
    launches0 = runner.launch_count
    st0 = runner.stats()
    sampler = ClockSampler(dev.index or 0)
    if rank == 0:
        sampler.start(); time.sleep(0.3)
    ms = timed_steps(D, stream, lambda: runner.run(U, Y, check_status=False), 0, steps) * steps
    clocks = sampler.stop() if rank == 0 else None
    launches = runner.launch_count - launches0
    st = runner.stats()
    bad = int((runner.status()[0] != 0).sum())
    it_tot, sol_tot, hom_tot, nc_tot, bad_tot = D.sumr([st["newton_iters"] - st0["newton_iters"], st["solves"] - st0["solves"],
                                                        st["homotopy_solves"] - st0["homotopy_solves"], st["not_converged"], bad])
    # ---- the only collective: final gather of the output shards (NCCL all-gather over NVLink)
    gather_ms = 0.0
    if world > 1:
        sharded.gather(Y); torch.cuda.synchronize(); D.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record(stream); yfull = sharded.gather(Y); g1.record(stream)
        torch.cuda.synchronize()
        gather_ms = D.maxr(g0.elapsed_time(g1))
        assert tuple(yfull.shape) == (C4_BATCH, n, 1)
        del yfull
    # ---- end to end: this rank's shard from / to pinned host buffers
    e2e = None
    if want_e2e:
        hu = torch.empty((count, n, 4), dtype=torch.float64, pin_memory=True)
        hy = torch.empty((count, n, 1), dtype=torch.float64, pin_memory=True)
        hu.copy_(U.cpu())
        del U
        torch.cuda.empty_cache()
        steps_e = max(1, min(steps, 2))
        runner.run_host_pinned(hu.data_ptr(), 4 * n, hy.data_ptr(), n)
        D.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter(); l0 = runner.launch_count
        for _ in range(steps_e):
            runner.run_host_pinned(hu.data_ptr(), 4 * n, hy.data_ptr(), n)
        torch.cuda.synchronize()
        dt = D.maxr(time.perf_counter() - t0)
        D.barrier()
        e2e = {"value": C4_BATCH * n * steps_e / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": int(C4_BATCH * n * 32),
               "d2h_bytes_per_step": int(C4_BATCH * n * 8), "steps": steps_e, "kernel_launches": int(runner.launch_count - l0),
               "checksum": float(hy[0, :, 0].abs().sum())}
        del hu, hy
    kernel_name = runner.kernel_name
    stored, cap = runner.cache_sizes()
    runner.close()
    torch.cuda.empty_cache()
    out = None
    if rank == 0:
        peak, peak_src = measured_peak()
        value = C4_BATCH * n * steps / (ms / 1e3) / 1e6
        kernel_ms = ms / steps
        achieved = C4_ALG_BYTES * (count * n / (kernel_ms / 1e3)) / 1e9
        spots = [37, 4095, 8000, 127]
        us = np.asfortranarray(c4_inputs_np(0, C4_BATCH, min(n, 8820))[spots].transpose(2, 1, 0))
        out = {"metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": steps, "warmup": warmup,
               "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic", "target": {"value": 100.0, "unit": "Msamples/s", "on": "8 GPUs (north_star)"},
               "config": {"workload": workload, "global_batch": C4_BATCH, "batch_per_gpu": count, "samples": n, "solver": SOLVER,
                          "parallelism": f"instances sharded over {world} GPU(s), no data-path collective; output all-gather timed separately",
                          "l2": "per-instance solver state lives on chip; the streams (2.9 GB in + 2.9 GB out per step in total) are touched once",
                          "kernel": kernel_name,
                          "state": f"the model state and the learnt solution cache persist from step to step: the timed steps are seconds "
                                   f"{warmup + 1}..{warmup + steps} of one continuous run from x = 0"},
               "clocks": clocks, "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                            "peak_source": peak_src, "algorithmic_bytes_per_launch": C4_ALG_BYTES * count * n, "kernel_ms": kernel_ms,
                            "note": "HBM fraction is the asked-for metric; this kernel is bound by the per-sample critical path of a "
                                    "13x13 pivoted LU inside one warp (profiles/)"},
               "gather": {"ms": gather_ms, "bytes": int(C4_BATCH * n * 8), "collective": "NCCL all_gather_into_tensor" if world > 1 else "none (1 GPU)",
                          "value_with_gather": C4_BATCH * n * steps / ((ms + gather_ms * steps) / 1e3) / 1e6},
               "newton": {"mean_iters": it_tot / max(sol_tot, 1), "homotopy_solves": int(hom_tot), "not_converged": int(nc_tot),
                          "instances_with_status": int(bad_tot), "stored_solutions_rank0": {"mean": float(stored.mean()), "max": int(stored.max()), "capacity": cap},
                          "note": "timed steps, all ranks"},
               "parity": parity_spot(model, len(spots), us, SOLVER, "4 instances of the sweep, the first 0.2 s from x = 0 (supply switching on)")}
        if e2e is not None:
            out["e2e"] = e2e
        if want_cpu:
            out["cpu_baseline"] = c4_cpu_baseline()
    return out


def run_config4(args, rank: int, world: int, local: int, emit):
    if args.impl == "reference":
        if rank != 0:
            return
        workload = ("examples/superover.jl (pots as inputs), batch=8192 with swept drive (128) x tone (64), level 1, "
                    "1 s of unit 1 kHz sine @ 44.1 kHz per instance (BASELINE.json configs[3])")
        vals, mss, info = [], [], None
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter(); info = c4_cpu_baseline(budget_s=6.0)
            if i >= args.warmup:
                vals.append(info["value"]); mss.append((time.perf_counter() - t0) * 1e3)
        v = float(np.mean(vals)) if vals else info["value"]
        info["value"] = v
        emit(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus,
                         "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(mss)) if mss else None,
                         "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                         "config": {"workload": workload + "; each step is a bounded sample of that sweep on the host cores"},
                         "cpu_baseline": info, "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                         "gpu_launches": 0}))
        return
    D = init_device(rank, world, local)
    out = config4_record(args, D, args.steps, args.warmup, not args.no_e2e, world == 1 and not args.no_cpu)
    if rank == 0:
        emit(json.dumps(out))
    finish(D)


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[4]: birdie.jl, batch = 32 768, 10 s of white noise, Newton-iteration histogram + HBM GB/s
C5_BATCH, C5_SECONDS = 32768, 10


def config5_record(args, D: Dist) -> dict:
    """One step = the 10 s of signal, as 10 launches of one second each: the input (white Gaussian noise, sigma 0.2 V,
    clipped to +-1 V) is generated on the device per time chunk, because 10 s of U + Y for 32 768 instances (231 GB)
    do not fit beside each other on one GPU; Y is overwritten chunk by chunk (checksummed)."""
    import torch
    from acme_jl_b200 import examples as ex
    from acme_jl_b200.distributed import ShardedBatchRunner
    dev, rank, world = D.dev, D.rank, D.world
    n = N_SAMPLES   # one second per launch
    model = ex.birdie(vol=0.8)
    sharded = ShardedBatchRunner(model, C5_BATCH, rank=rank, world=world, solver=SOLVER, kernel=args.kernel)
    runner, first, count = sharded.runner, sharded.first, sharded.count
    g = torch.Generator(device=dev); g.manual_seed(0xACE5EED + rank)
    Y = torch.empty((count, n, 1), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)

    def noise():
        return (0.2 * torch.randn((count, n, 1), generator=g, device=dev, dtype=torch.float64)).clamp_(-1, 1)

    for _ in range(3):   # warm-up: three seconds of signal (three launches)
        runner.run(noise(), Y, check_status=False)
    torch.cuda.synchronize()
    st0 = runner.stats()
    l0 = runner.launch_count
    tot_ms, chk = 0.0, 0.0
    D.barrier()
    for _ in range(C5_SECONDS):
        U = noise()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); runner.run(U, Y, check_status=False); e1.record(stream)
        torch.cuda.synchronize()
        tot_ms += e0.elapsed_time(e1)
        chk += float(Y[0, :, 0].abs().sum())
    D.barrier()
    tot_ms = D.maxr(tot_ms)
    st = runner.stats()
    bad = int((runner.status()[0] != 0).sum())
    hist = D.sumr([a - b for a, b in zip(st["iter_hist"], st0["iter_hist"])])
    it_tot, sol_tot, hom_tot, nc_tot, bad_tot = D.sumr([st["newton_iters"] - st0["newton_iters"], st["solves"] - st0["solves"],
                                                        st["homotopy_solves"] - st0["homotopy_solves"], st["not_converged"], bad])
    stored, cap = runner.cache_sizes()
    launches = runner.launch_count - l0
    kernel_name = runner.kernel_name
    Us = noise()[:4].cpu().numpy()   # spot-check input (host copy for the oracle)
    runner.close()
    del Y
    torch.cuda.empty_cache()
    out = None
    if rank == 0:
        peak, peak_src = measured_peak()
        total = C5_BATCH * n * C5_SECONDS
        rate_gpu = count * n * C5_SECONDS / (tot_ms / 1e3)
        cap_ncu = ncu_capture("r2/traffic_birdie.json")
        out = {"workload": "examples/birdie.jl (vol = 0.8), batch=32768 independent white-noise inputs (sigma 0.2 V, clipped to +-1 V), "
                           "10 s @ 44.1 kHz, sharded over the GPUs (BASELINE.json configs[4]); input generated on the device per "
                           "1-s chunk (10 s of U + Y = 231 GB do not fit one GPU), Y overwritten per chunk",
               "value": total / (tot_ms / 1e3) / 1e6, "unit": "Msamples/s", "ms_per_step": tot_ms, "steps": 1, "warmup": "3 launches (3 s of signal)",
               "scaling": "strong", "global_batch": C5_BATCH, "batch_per_gpu": count, "samples": n * C5_SECONDS, "kernel": kernel_name,
               "gpu_launches": int(launches), "solver": SOLVER,
               "roofline": {"bound": "hbm", "achieved": ALG_BYTES_PER_SAMPLE * rate_gpu / 1e9, "peak": peak, "unit": "GB/s",
                            "frac": ALG_BYTES_PER_SAMPLE * rate_gpu / 1e9 / peak, "peak_source": peak_src,
                            "traffic": cap_ncu.get("dram_bytes_per_launch") if cap_ncu else None,
                            "algorithmic_bytes_per_launch": ALG_BYTES_PER_SAMPLE * count * n, "kernel_ms": tot_ms / C5_SECONDS,
                            "note": "latency / FP64-issue bound (Newton on a 2x2 system with a Gummel-Poon BJT), not HBM"},
               "newton": {"mean_iters": it_tot / max(sol_tot, 1), "hist_bins_1_to_31_then_32plus": [int(h) for h in hist],
                          "homotopy_solves": int(hom_tot), "solves": int(sol_tot), "not_converged": int(nc_tot), "instances_with_status": int(bad_tot),
                          "stored_solutions_rank0": {"mean": float(stored.mean()), "max": int(stored.max()), "capacity": cap}},
               "checksum": chk,
               "parity": parity_spot(model, 4, np.asfortranarray(Us.transpose(2, 1, 0)[:, :min(n, 22050)]), SOLVER,
                                     "4 instances, 0.5 s of the same kind of noise from x = 0")}
    return out


def init_device(rank: int, world: int, local: int) -> Dist:
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    return Dist(rank, world, dev)


def finish(D: Dist):
    if D.world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


def _stdout_to_stderr():
    """Everything except the final JSON line goes to stderr (NCCL prints its version banner on
    stdout during init); returns a function that prints one line on the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)

    def emit(line: str):
        sys.stdout.flush()
        os.write(real, (line + "\n").encode())

    return emit


def main():
    global N_SAMPLES
    emit = _stdout_to_stderr()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="instances per GPU (default: the BASELINE config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--layout", default="instance", choices=["instance", "sample"],
                    help="stream layout of config 2: instance-major (nu,N,B) = the reference's per-instance blocks (default, "
                         "the measured headline) or sample-major (nu,B,N) (ACMEB200_SAMPLE_MAJOR, DESIGN.md 4.1b)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4],
                    help="BASELINE.json configs index of the TOP-LEVEL record: 2 = diode clipper sweep (the headline, default, weak "
                         "scaling); 4 = superover B=8192 sharded over the GPUs (strong scaling, final output gather timed separately)")
    ap.add_argument("--sub", default="1,3,4,5",
                    help="further BASELINE configs measured in the same run and reported as sub-records config1/config3/config4/"
                         "config5 of the one JSON line ('none' to skip): 1 = one instance on one CPU core (oracle), 3 = Sallen-Key "
                         "65536 @ 96 kHz (HBM-bound), 4 = superover 8192 sharded over the GPUs + NCCL gather, 5 = birdie 32768, 10 s of noise")
    ap.add_argument("--sub-steps", type=int, default=3, help="timed steps of the sub-records (3 warm-up steps each)")
    ap.add_argument("--c3-batch", type=int, default=C3_BATCH, help="config 3 instances per GPU (smaller values: smoke runs only)")
    ap.add_argument("--samples", type=int, default=N_SAMPLES, help="samples per instance (default 44100 = 1 s; "
                    "smaller values are for profiling under ncu only, such a line is not a bench value)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    N_SAMPLES = args.samples
    subs = set() if args.sub in ("none", "") else {int(c) for c in args.sub.split(",")}

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.config == 4:
        run_config4(args, rank, world, local, emit)
        return
    if args.impl == "reference":
        run_reference(args, rank, world, emit)
        return

    # host-side model derivation of config 3 first: its worker pool forks, which must happen before this process
    # holds a CUDA context (rank 0 derives, the matrices are broadcast)
    pre3, sub_errors = None, {}
    if 3 in subs and rank == 0:
        try:
            pre3 = c3_derive(args.c3_batch)
        except Exception as e:  # the headline must not die with a sub-record
            sub_errors["config3"] = f"{type(e).__name__}: {e}"

    import torch
    import torch.distributed as dist
    D = init_device(rank, world, local)
    dev = D.dev

    from acme_jl_b200 import BatchRunner, examples as ex
    from acme_jl_b200.distributed import ShardedBatchRunner
    Bper = args.batch
    model = ex.diodeclipper()
    # one descriptor for the global batch; every rank uploads only its contiguous shard
    Pglobal = sweep_params(Bper * world, 0, Bper * world)
    sharded = ShardedBatchRunner(model, Bper * world, rank=rank, world=world, params=[Pglobal], solver=SOLVER,
                                 kernel=args.kernel)
    assert sharded.count == Bper and sharded.first == rank * Bper
    runner = sharded.runner
    P = Pglobal[:, rank * Bper:(rank + 1) * Bper]

    row = torch.from_numpy(sine_row()).to(dev)
    smaj = args.layout == "sample"
    if smaj:  # (N, B, nu): one time step of the whole shard contiguous
        U = row.reshape(N_SAMPLES, 1, 1).expand(N_SAMPLES, Bper, 1).contiguous()
        Y = torch.empty((N_SAMPLES, Bper, 1), dtype=torch.float64, device=dev)
    else:     # (B, N, nu): per-instance streams in HBM
        U = row.reshape(1, N_SAMPLES, 1).expand(Bper, N_SAMPLES, 1).contiguous()
        Y = torch.empty((Bper, N_SAMPLES, 1), dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream(dev)
    barrier = D.barrier

    for _ in range(args.warmup):
        runner.run(U, Y, check_status=False, layout=args.layout)
    torch.cuda.synchronize()
    launches0 = runner.launch_count
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        runner.run(U, Y, check_status=False, layout=args.layout)
    e1.record(stream)
    torch.cuda.synchronize(); barrier()
    ms = D.maxr(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    launches = runner.launch_count - launches0
    st = runner.stats()
    status_bad = int((runner.status()[0] != 0).sum())
    samples_step = Bper * world * N_SAMPLES
    value = samples_step * args.steps / (ms / 1e3) / 1e6
    kernel_ms = ms / args.steps  # one kernel launch per step: launch duration == step duration

    # ---- end to end: pinned host buffers, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        import psutil
        need = 2 * Bper * N_SAMPLES * 8
        avail = psutil.virtual_memory().available / max(world, 1)
        Be = Bper
        while Be > 1024 and need * (Be / Bper) * 1.6 > avail:
            Be //= 2
        del U, Y
        torch.cuda.empty_cache()
        r2 = runner if Be == Bper else BatchRunner(model, Be, params=[P[:, :Be]], solver=SOLVER, kernel=args.kernel)
        hshape = (N_SAMPLES, Be, 1) if smaj else (Be, N_SAMPLES, 1)
        hu = torch.empty(hshape, dtype=torch.float64, pin_memory=True)
        hy = torch.empty(hshape, dtype=torch.float64, pin_memory=True)
        hu.copy_(row.cpu().reshape((N_SAMPLES, 1, 1) if smaj else (1, N_SAMPLES, 1)).expand(*hshape))
        e2e_steps = max(1, min(args.steps, 3))
        hstride = Be if smaj else N_SAMPLES  # doubles between samples / between instances
        r2.run_host_pinned(hu.data_ptr(), hstride, hy.data_ptr(), N_SAMPLES, layout=args.layout)  # warm-up (allocates staging)
        barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        l0 = r2.launch_count
        for _ in range(e2e_steps):
            r2.run_host_pinned(hu.data_ptr(), hstride, hy.data_ptr(), N_SAMPLES, layout=args.layout)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        le = r2.launch_count - l0
        barrier()
        dt = D.maxr(dt)
        e2e = {"value": Be * world * N_SAMPLES * e2e_steps / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(Be * N_SAMPLES * 8), "d2h_bytes_per_step": int(Be * N_SAMPLES * 8),
               "batch_per_gpu": Be, "steps": e2e_steps, "kernel_launches": int(le),
               "host_copy_GBs_per_direction": Be * world * N_SAMPLES * 8 * e2e_steps / dt / 1e9,
               "checksum": float((hy[:, 0, 0] if smaj else hy[0, :, 0]).abs().sum())}
        if r2 is not runner:
            r2.close()
        del hu, hy
        try:   # the bound, measured: how fast this box's host side moves the same bytes without any kernel
            e2e["host_copy_ceiling"] = host_copy_ceiling(D)
        except Exception as e:
            e2e["host_copy_ceiling"] = {"error": f"{type(e).__name__}: {e}"}
    else:
        del U, Y
    kernel_name = runner.kernel_name
    runner.close()
    torch.cuda.empty_cache()

    out = None
    if rank == 0:
        from acme_jl_b200._lib import measure_fp64_peak
        fp64_peak = measure_fp64_peak()
        peak, peak_src = measured_peak()
        per_gpu_rate = Bper * N_SAMPLES / (kernel_ms / 1e3)
        achieved = ALG_BYTES_PER_SAMPLE * per_gpu_rate / 1e9
        cap = ncu_capture("r2/traffic_clipper.json")   # refreshed by tools/ncu_refresh.py; None when stale
        traffic = cap["dram_bytes_per_sample"] * Bper * N_SAMPLES if cap else None
        flops = cap["fp64_flops_per_sample"] if cap else None
        spots = sorted({i % (Bper * world) for i in (0, 255, 256 * 255, 65535, 31337, 4242, 12345, 54321)})
        out = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "examples/diodeclipper.jl, batch=65536 instances/GPU with swept Is/eta (256x256), "
                                   "1 s of unit 1 kHz sine @ 44.1 kHz per instance (BASELINE.json configs[1])",
                       "batch_per_gpu": Bper, "global_batch": Bper * world, "samples": N_SAMPLES, "solver": SOLVER,
                       "parallelism": f"instances sharded over {world} GPU(s), no data-path collective",
                       "l2": "inputs larger than L2 (23 GB U + 23 GB Y per GPU per step)",
                       "kernel": kernel_name,
                       "layout": "sample-major (nu,B,N) streams" if smaj else "instance-major (nu,N,B) streams"},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "traffic_source": (f"ncu dram__bytes of {cap['capture']}, per sample x the samples of this launch; kernel sources "
                                            f"unchanged since (stamp {cap['src_stamp']})") if cap else
                                           "no ncu capture of the kernel sources in this tree (profiles/r2/traffic_clipper.json missing or stale)",
                         "note": "HBM fraction is the asked-for metric; the kernel is FP64-pipe / issue bound (see DESIGN.md)",
                         "algorithmic_bytes_per_launch": ALG_BYTES_PER_SAMPLE * Bper * N_SAMPLES,
                         "kernel_ms": kernel_ms},
            "fp64_pipe": {"measured_dfma_peak_tflops": fp64_peak,
                          "note": "DFMA microbenchmark in this run (acmeb200_measure_fp64_peak)"},
            "newton": {"mean_iters": st["newton_iters"] / max(st["solves"], 1), "hist_1_to_8": st["iter_hist"][:8],
                       "homotopy_solves": st["homotopy_solves"], "not_converged": st["not_converged"],
                       "instances_with_status": status_bad},
            "parity": parity_spot(model, len(spots), sine_row().reshape(1, -1), SOLVER,
                                  f"{len(spots)} instances spread over the (Is, eta) grid, the full second", params=[Pglobal[:, spots]]),
        }
        if flops:
            # the roofline that actually binds this kernel (FP64 CUDA-core pipe; dependent-issue latency keeps it
            # from the peak): executed FP64 flops per sample x samples/s against the DFMA peak measured in this run
            out["roofline_fp64"] = {"bound": "fp64", "achieved": flops * per_gpu_rate / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                                    "frac": flops * per_gpu_rate / 1e12 / fp64_peak if fp64_peak else None, "flops_per_sample": flops,
                                    "source": f"ncu instruction mix of {cap['capture']}"}
        if e2e is not None:
            out["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline()

    # ---- the other BASELINE configs, each a sub-record of the same line (a failure is recorded, never fatal)
    def sub(name, fn):
        try:
            r = fn()
        except Exception as e:
            import traceback
            traceback.print_exc()
            r = {"error": f"{type(e).__name__}: {e}"} if rank == 0 else None
            try:
                torch.cuda.empty_cache()
            except Exception:
                pass
        if rank == 0:
            out[name] = r

    if 3 in subs:
        ok = D.sumr([0.0 if (rank == 0 and pre3 is None) else 1.0])[0] == world   # every rank must agree to run it
        if ok:
            sub("config3", lambda: config3_record(args, D, pre3))
        elif rank == 0:
            out["config3"] = {"error": sub_errors.get("config3", "derivation failed")}
    if 4 in subs:
        sub("config4", lambda: config4_record(args, D, args.sub_steps, 3, not args.no_e2e, world == 1 and not args.no_cpu))
    if 5 in subs:
        sub("config5", lambda: config5_record(args, D))
    if 1 in subs and rank == 0 and not args.no_cpu:
        try:
            out["config1"] = config1_record()
        except Exception as e:
            out["config1"] = {"error": f"{type(e).__name__}: {e}"}
    if rank == 0:
        emit(json.dumps(out))
    finish(D)


if __name__ == "__main__":
    main()
