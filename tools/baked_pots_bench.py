"""alternative reading of config 4: superover with baked potentiometers, per-instance matrices (derive_sweep)"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import acme_jl_b200 as A
from acme_jl_b200 import BatchRunner, examples as ex
def build(d, t): return ex.superover(d, t, 1.0)
B = int(os.environ.get("KB_B", 1024)); N = int(os.environ.get("KB_N", 2205)); D = int(os.environ.get("KB_DISTINCT", 64))
pts = [((k % 8 + 0.5) / 8, (k // 8 % 8 + 0.5) / 8) for k in range(D)]
t0 = time.perf_counter(); base, kw, _ = A.derive_sweep(build, pts, chunk=2); dt = time.perf_counter() - t0
rep = B // D
kw2 = {"overrides": {k: np.asfortranarray(np.tile(v, (1,) * (v.ndim - 1) + (rep,))) for k, v in kw["overrides"].items()},
       "init_z": [np.asfortranarray(np.tile(z, (1, rep))) for z in kw["init_z"]]}
HC = "HomotopySolver{CachingSolver{SimpleSolver}}"
dev = torch.device("cuda", 0)
U = torch.sin(2 * np.pi * 1000 / 44100 * torch.arange(N, device=dev, dtype=torch.float64)).reshape(1, N, 1).expand(B, N, 1).contiguous()
Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
for kernel in os.environ.get("KB_KERNELS", "auto").split(","):
    r = BatchRunner(base, B, solver=HC, kernel=kernel, **kw2)
    r.run(U, Y, check_status=False); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.run(U, Y, check_status=False); e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)
    st = r.stats()
    print(json.dumps(dict(derive_s=round(dt, 2), distinct=D, B=B, N=N, kernel=r.kernel_name[:40], ms=round(ms, 2), Msamples_s=round(B * N / ms / 1e3, 2),
                          iters=round(st["newton_iters"] / st["solves"], 3), bad=int((r.status()[0] != 0).sum()), chk=float(Y[B // 3, -1, 0]))))
    r.close()
