"""small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import acme_jl_b200 as A
from acme_jl_b200 import BatchRunner, examples as ex
HC = "HomotopySolver{CachingSolver{SimpleSolver}}"
sine = lambda n: np.sin(2 * np.pi * 1000 / 44100 * np.arange(n)).reshape(1, -1)
def so_inputs(B, N):
    u = np.zeros((4, N, B), order="F"); u[0] = sine(N)[0][:, None]
    u[1] = ((np.arange(B) * 37 % 128 + 0.5) / 128)[None, :]; u[2] = 0.5; u[3] = 1.0
    return u
out = []
ONLY = os.environ.get("SAN_ONLY", "")  # "rows": just the rows kernel
SMALL = os.environ.get("SAN_SMALL", "") == "1"  # the -m gpu test: short runs (the sanitizer slows the kernels ~50x)
# rows kernel: 1-warp CTAs, the 4-warp build (ragged last CTA), both solvers
for B, N in (((3, 400), (1301, 40)) if not SMALL else ((2, 260), (1190, 6))) if ONLY != "rows" else ((3, 150), (1190, 12)):
    r = BatchRunner(ex.superover(), B, solver=HC); y = r.run(so_inputs(B, N)); out.append((r.kernel_name[:20], B, float(np.abs(y).sum()))); r.close()
# rows kernel, per-instance matrices (baked pots)
base, kw, B = A.derive_sweep(lambda d, t: ex.superover(d, t, 1.0), [(0.2 + 0.3 * k, 0.5) for k in range(3)], workers=1)
r = BatchRunner(base, B, solver=HC, **kw); y = r.run(sine(120 if SMALL else 300)); out.append((r.kernel_name[:20], B, float(np.abs(y).sum()))); r.close()
if ONLY == "rows":
    for o in out: print(o)
    sys.exit(0)
# cooperative and generic kernels on the same model
for k in ("coop", "generic"):
    r = BatchRunner(ex.superover(), 5, solver=HC, kernel=k); y = r.run(so_inputs(5, 60 if SMALL else 120)); out.append((r.kernel_name[:20], 5, float(np.abs(y).sum()))); r.close()
# thread-per-instance kernels: clipper (partial warp, partial tile), linear with per-instance matrices, birdie + noise
P = np.vstack([np.full(100, 1e-15), np.linspace(1, 2, 100), np.full(100, 1.8e-15), np.linspace(1, 2, 100)])
r = BatchRunner(ex.diodeclipper(), 100, params=[P], solver=HC); y = r.run(np.asfortranarray(np.repeat(sine(203)[:, :, None], 100, axis=2))); out.append((r.kernel_name[:20], 100, float(np.abs(y).sum()))); r.close()
base, kw, B = A.derive_sweep(lambda R: ex.sallenkey(fs=96000, r1=R, r2=R), [1e3 * (1 + k) for k in range(70)], workers=1)
r = BatchRunner(base, B, **kw); y = r.run(np.asfortranarray(np.repeat(sine(101)[:, :, None], B, axis=2))); out.append((r.kernel_name[:20], B, float(np.abs(y).sum()))); r.close()
rng = np.random.default_rng(1)
r = BatchRunner(ex.birdie(vol=0.8), 40, solver=HC); y = r.run(np.asfortranarray(np.clip(0.2 * rng.standard_normal((1, 600, 40)), -1, 1))); out.append(("stored", int(r.cache_sizes()[0].max()), 0.0)); out.append((r.kernel_name[:20], 40, float(np.abs(y).sum()))); r.close()
for o in out: print(o)
