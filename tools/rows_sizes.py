"""rows kernel throughput at the per-GPU batch sizes of the 1/2/4/8-GPU runs of config 4"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
HC = "HomotopySolver{CachingSolver{SimpleSolver}}"
m = ex.superover(); dev = torch.device("cuda", 0)
N = 2205
out = {"lib": os.path.basename(os.environ.get("ACMEB200_LIB", "default"))}
for B in (2048, 4096, 8192):
    U = torch.zeros((B, N, 4), dtype=torch.float64, device=dev)
    U[:, :, 0] = torch.sin(2 * np.pi * 1000 / 44100 * torch.arange(N, device=dev, dtype=torch.float64))[None, :]
    k = torch.arange(B, device=dev)
    U[:, :, 1] = (((k % 128) + 0.5) / 128)[:, None]; U[:, :, 2] = (((k // 128) % 64 + 0.5) / 64)[:, None]; U[:, :, 3] = 1.0
    Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
    r = BatchRunner(m, B, solver=HC, kernel="rows")
    r.run(U, Y, check_status=False); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.run(U, Y, check_status=False); e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)
    out[f"B{B}"] = round(B * N / ms / 1e3, 2)
    r.close(); del U, Y
print(json.dumps(out))
