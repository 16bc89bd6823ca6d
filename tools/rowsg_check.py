"""grouped rows kernel variants (ACMEB200_ROWS_VARIANT) against the default rows kernel: bit equality, statistics, timing"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
H, HC = "HomotopySolver{SimpleSolver}", "HomotopySolver{CachingSolver{SimpleSolver}}"
m = ex.superover()
dev = torch.device("cuda", 0)
def inputs(B, N):
    u = np.zeros((B, N, 4))
    u[:, :, 0] = np.sin(2 * np.pi * 1000 / 44100 * np.arange(N))[None, :]
    u[:, :, 1] = ((np.arange(B) * 37 % 128) + 0.5)[:, None] / 128
    u[:, :, 2] = ((np.arange(B) * 11 % 64) + 0.5)[:, None] / 64
    u[:, :, 3] = 1.0
    return torch.from_numpy(u).to(dev)
def run(variant, B, N, solver, reps=1):
    if variant: os.environ["ACMEB200_ROWS_VARIANT"] = variant
    else: os.environ.pop("ACMEB200_ROWS_VARIANT", None)
    r = BatchRunner(m, B, solver=solver, kernel="rows")
    U = inputs(B, N); Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
    r.run(U, Y, check_status=False); torch.cuda.synchronize()
    st = r.stats(); y1 = Y.clone()
    ms = None
    if reps:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.run(U, Y, check_status=False); e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)
    bad = int((r.status()[0] != 0).sum()); r.close()
    return y1, st, ms, bad
variants = os.environ.get("RG_VARIANTS", "g1w1,g2w1,g2w4").split(",")
if os.environ.get("RG_PARITY", "1") == "1":
    for solver in (H, HC):
        for B, N in ((7, 1500), (64, 600)):
            yref, sref, _, _ = run("", B, N, solver, reps=0)
            for v in variants:
                y, s, _, bad = run(v, B, N, solver, reps=0)
                print(solver[:22], B, N, v, "equal", bool(torch.equal(y, yref)), "maxdiff", float((y - yref).abs().max()), "hist_equal", s["iter_hist"] == sref["iter_hist"],
                      "hom", s["homotopy_solves"], sref["homotopy_solves"], "samples", s["samples"], sref["samples"], "bad", bad, flush=True)
for B, N in ((1024, 4410), (8192, 2205)):
    for v in [""] + variants:
        y, s, ms, bad = run(v, B, N, HC)
        print(json.dumps(dict(B=B, N=N, variant=v or "default", ms=round(ms, 2), Msamples_s=round(B * N / ms / 1e3, 2), iters=round(s["newton_iters"] / s["solves"], 3), bad=bad, chk=float(y[B // 3, -1, 0]))), flush=True)
