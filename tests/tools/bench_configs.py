"""All five BASELINE.json configs on one GPU (kernel-only, streams resident in HBM), with the CPU
oracle on a bounded sample beside each.  Writes a markdown table (stdout) -- the numbers quoted in
DESIGN.md section 7 / profiles/configs_r1.md.  Not the driver's bench (that is bench.py = config 2).
CONFIGS=3,5 restricts the run; NO_CPU=1 skips the oracle columns.  Configs 2 and 3 (thread-per-instance kernels) add a
second row timed on sample-major streams (ACMEB200_SAMPLE_MAJOR)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
from oracle.oracle import OracleModel
import bench

dev = torch.device("cuda", 0)
CORES = bench.host_cores()
ONLY = set(int(c) for c in os.environ.get("CONFIGS", "2,3,4,5").split(","))
NO_CPU = os.environ.get("NO_CPU", "0") == "1"
rows, notes = [], []


def timed(runner, U, Y, steps=2, warm=1, **kw):
    for _ in range(warm): runner.run(U, Y, check_status=False, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): runner.run(U, Y, check_status=False, **kw)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def cpu(model, b, u, **kw):
    if NO_CPU:
        return float("nan"), float("nan")
    o = OracleModel(model, b, **kw)
    t0 = time.perf_counter(); o.run(u, threads=CORES); dt = time.perf_counter() - t0
    st = o.stats()
    return b * u.shape[1] / dt / 1e6, st["newton_iters"] / max(st["solves"], 1)


def sine(N, fs): return np.sin(2 * np.pi * 1000 / fs * np.arange(N))


def cfg2():  # diode clipper B=65536, 1 s @ 44.1 kHz
    B, N = 65536, 44100
    m = ex.diodeclipper(); P = bench.sweep_params(B, 0, B)
    r = BatchRunner(m, B, params=[P])
    U = torch.from_numpy(sine(N, 44100)).to(dev).reshape(1, N, 1).expand(B, N, 1).contiguous(); Y = torch.empty_like(U)
    ms = timed(r, U, Y); st = r.stats()
    c, ci = cpu(m, 32 * CORES, sine(N, 44100).reshape(1, -1), params=[P[:, ::B // (32 * CORES)][:, :32 * CORES]])
    rows.append(("2 diode clipper, B=65536 swept Is/eta, 1 s @44.1 kHz", r.kernel_name, B * N / ms / 1e3, 16 * B * N / ms / 1e6, st["newton_iters"] / st["solves"], c, ci))
    r.reset()
    Us = U.transpose(0, 1).contiguous(); Ys = torch.empty_like(Us)   # (N, B, nu): sample-major streams
    ms = timed(r, Us, Ys, layout="sample"); st = r.stats()
    assert torch.equal(Ys.transpose(0, 1), Y), "sample-major result differs from the default layout"
    rows.append(("2 (same, sample-major streams)", r.kernel_name, B * N / ms / 1e3, 16 * B * N / ms / 1e6, st["newton_iters"] / st["solves"], float("nan"), float("nan")))
    r.close()


def cfg3():  # Sallen-Key, per-instance matrices, 1 s @ 96 kHz (256 distinct (R, kappa) pairs tiled to 65536)
    import acme_jl_b200 as A
    B, N = 65536, 96000
    pts = [(10 ** (3 + 2 * (k % 16) / 15), 10 ** (1.3 * (k // 16) / 15)) for k in range(256)]
    t0 = time.perf_counter()
    base, kw, _ = A.derive_sweep(lambda R, kap: ex.sallenkey(fs=96000, r1=R, r2=R, c1=10e-9 * kap, c2=10e-9 / kap), pts)
    notes.append(f"config 3: 256 models derived on the host in {time.perf_counter() - t0:.2f} s (derive_sweep, {CORES} cores)")
    ov = {k: np.tile(v, (1,) * (v.ndim - 1) + (B // 256,)) for k, v in kw["overrides"].items()}
    r = BatchRunner(base, B, overrides=ov)
    U = torch.from_numpy(sine(N, 96000)).to(dev).reshape(1, N, 1).expand(B, N, 1).contiguous(); Y = torch.empty_like(U)
    ms = timed(r, U, Y)
    c, ci = cpu(base, 64 * CORES, sine(N, 96000).reshape(1, -1), overrides={k: v[..., :64 * CORES] for k, v in ov.items()})
    rows.append(("3 Sallen-Key, B=65536 per-instance matrices (256 distinct R/C pairs tiled), 1 s @96 kHz", r.kernel_name, B * N / ms / 1e3, 16 * B * N / ms / 1e6, 0.0, c, ci))
    r.reset()
    Yk = Y[:, ::97].clone(); del Y                                    # 4 x 50 GB would not fit: keep every 97th sample
    Us = U.transpose(0, 1).contiguous(); del U                        # (N, B, nu): sample-major streams
    Ys = torch.empty_like(Us)
    ms = timed(r, Us, Ys, layout="sample")
    assert torch.equal(Ys[::97].transpose(0, 1), Yk), "sample-major result differs from the default layout"
    rows.append(("3 (same, sample-major streams)", r.kernel_name, B * N / ms / 1e3, 16 * B * N / ms / 1e6, 0.0, float("nan"), float("nan")))
    r.close()


def cfg4():  # superover, pots as inputs, 1 s @ 44.1 kHz; 8192 on one GPU and 1024 = the per-GPU share of 8; seconds 1 and 3
    m = ex.superover()
    c = ci = float("nan")
    for B in (8192, 1024):
        N = 44100
        r = BatchRunner(m, B)
        U = torch.from_numpy(bench.c4_inputs_np(0, B, N)).to(dev)
        Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
        prev = dict(newton_iters=0, solves=0)
        for sec in (1, 2, 3):
            ms = timed(r, U, Y, steps=1, warm=0); st = r.stats(); bad = int((r.status()[0] != 0).sum())
            it = (st["newton_iters"] - prev["newton_iters"]) / (st["solves"] - prev["solves"]); prev = st
            if sec == 2:
                continue
            if B == 8192 and sec == 1:
                b = 2 * CORES
                uc = np.asfortranarray(bench.c4_inputs_np(0, 8192, 8820)[:: 8192 // b][:b].transpose(2, 1, 0))
                c, ci = cpu(m, b, uc)
            rows.append((f"4 superover (pots as inputs), B={B}, 1 s @44.1 kHz, second {sec} of the run from x=0, status!=0: {bad}", r.kernel_name, B * N / ms / 1e3, 40 * B * N / ms / 1e6, it, c, ci))
        del U, Y; r.close(); torch.cuda.empty_cache()


def cfg5():  # birdie(vol=0.8), B=32768, 10 s of clipped white noise, chunked in time (U+Y = 231 GB > 180 GB)
    B, N, chunks = 32768, 44100, 10
    m = ex.birdie(vol=0.8)
    r = BatchRunner(m, B)
    g = torch.Generator(device=dev); g.manual_seed(0xACE5EED)
    Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
    tot_ms = 0.0
    for cidx in range(chunks):
        U = (0.2 * torch.randn((B, N, 1), generator=g, device=dev, dtype=torch.float64)).clamp_(-1, 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.run(U, Y, check_status=False); e1.record(); torch.cuda.synchronize()
        tot_ms += e0.elapsed_time(e1)
    st = r.stats(); bad = int((r.status()[0] != 0).sum())
    rng = np.random.default_rng(1); b = 4 * CORES
    c, ci = cpu(m, b, np.asfortranarray(np.clip(0.2 * rng.standard_normal((1, 44100, b)), -1, 1)))
    rows.append((f"5 birdie(vol=0.8), B=32768, 10 s clipped white noise (sigma 0.2), 10 time chunks, noise generated on device, status!=0: {bad}", r.kernel_name, B * N * chunks / tot_ms / 1e3, 16 * B * N * chunks / tot_ms / 1e6, st["newton_iters"] / st["solves"], c, ci))
    notes.append(f"config 5 Newton-iteration histogram (bins 1..31, 32+): {st['iter_hist']} homotopy solves: {st['homotopy_solves']} of {st['solves']}; "
                 f"stored solutions per instance: mean {r.cache_sizes()[0].mean():.1f}")
    r.close()


for k, f in ((2, cfg2), (3, cfg3), (4, cfg4), (5, cfg5)):
    if k in ONLY:
        f(); torch.cuda.empty_cache()
print("| config | kernel | Msamples/s (1 B200, HBM-resident) | algorithmic GB/s | Newton iters/solve | CPU oracle Msamples/s (%d cores) | CPU iters/solve |" % CORES)
print("|---|---|---|---|---|---|---|")
for row in rows:
    print("| %s | `%s` | %.1f | %.1f | %.2f | %.3f | %.2f |" % row)
print()
for n in notes:
    print(n)
