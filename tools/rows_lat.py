"""per-sample latency of the rows kernel vs batch size (how much do co-resident warps slow each other down?)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
HC = "HomotopySolver{CachingSolver{SimpleSolver}}"
m = ex.superover(); dev = torch.device("cuda", 0)
N = 4410
for B in (1, 148, 296, 592, 1024, 2048, 2368):
    u = np.zeros((B, N, 4)); u[:, :, 0] = np.sin(2 * np.pi * 1000 / 44100 * np.arange(N))[None, :]
    u[:, :, 1] = ((np.arange(B) * 37 % 128) + 0.5)[:, None] / 128; u[:, :, 2] = ((np.arange(B) * 11 % 64) + 0.5)[:, None] / 64; u[:, :, 3] = 1.0
    U = torch.from_numpy(u).to(dev); Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
    r = BatchRunner(m, B, solver=HC, kernel=os.environ.get("KB_KERNEL", "rows"))
    r.run(U, Y, check_status=False); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.run(U, Y, check_status=False); e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)
    st = r.stats()
    print(json.dumps(dict(B=B, ms=round(ms, 2), us_per_sample=round(ms * 1e3 / N, 2), kcycles_per_sample=round(ms * 1e3 / N * 1.965, 1), Msamples_s=round(B * N / ms / 1e3, 2), iters=round(st["newton_iters"] / st["solves"], 2))), flush=True)
    r.close()
