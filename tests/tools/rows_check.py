"""quick GPU check of the rows-in-registers kernel: parity against the oracle and the cooperative
kernel on superover, iteration statistics, timing at B=1024 and B=8192"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
from oracle.oracle import OracleModel

H, HC = "HomotopySolver{SimpleSolver}", "HomotopySolver{CachingSolver{SimpleSolver}}"
m = ex.superover()


def inputs(B, N):
    u = np.zeros((4, N, B), order="F")
    u[0] = np.sin(2 * np.pi * 1000 / 44100 * np.arange(N))[:, None]
    u[1] = ((np.arange(B) % 128) + 0.5)[None, :] / 128
    u[2] = ((np.arange(B) // 128 % 64) + 0.5)[None, :] / 64
    u[3] = 1.0
    return u


if os.environ.get("RC_PARITY", "1") == "1":
    B, N = 16, 1500
    u = inputs(B, N)
    u[1] = (np.arange(B) % 4 + 0.5)[None, :] / 4
    u[2] = (np.arange(B) // 4 + 0.5)[None, :] / 4
    for solver in (H, HC):
        o = OracleModel(m, B, solver=solver)
        yref = o.run(u, threads=0)
        out = {}
        for k in ("rows", "coop"):
            r = BatchRunner(m, B, solver=solver, kernel=k)
            y = r.run(u)
            st = r.stats()
            peak = np.abs(yref).max()
            err = np.abs(y - yref) / np.maximum(np.abs(yref), 1e-3 * peak)
            out[k] = dict(name=r.kernel_name[:24], maxrel=float(np.nanmax(err)), nan=int(np.isnan(y).sum()),
                          iters=st["newton_iters"] / max(st["solves"], 1), hom=st["homotopy_solves"], nc=st["not_converged"],
                          bad=int((r.status()[0] != 0).sum()), hist=st["iter_hist"][:10])
            # state persistence / kernel switch: a second call continues
            r.close()
        so = o.stats()
        print(solver, json.dumps(out), "oracle iters", so["newton_iters"] / so["solves"], "hom", so["homotopy_solves"], so["iter_hist"][:10], flush=True)

dev = torch.device("cuda", 0)
for B, N in ((1024, 4410), (8192, 2205)):
    u = torch.from_numpy(np.ascontiguousarray(inputs(B, N).transpose(2, 1, 0))).to(dev)  # (B, N, nu)
    for k in ("rows", "coop"):
        r = BatchRunner(m, B, solver=HC, kernel=k)
        Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
        r.run(u, Y, check_status=False); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.run(u, Y, check_status=False); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        st = r.stats()
        print(json.dumps(dict(B=B, N=N, kernel=r.kernel_name[:28], ms=round(ms, 2), Msamples_s=round(B * N / ms / 1e3, 2),
                              iters=round(st["newton_iters"] / max(st["solves"], 1), 3), hom=st["homotopy_solves"],
                              bad=int((r.status()[0] != 0).sum()), chk=float(Y[B // 3, -1, 0]))), flush=True)
        r.close()
