import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
import bench
B = int(os.environ.get("KB_B", 65536)); N = int(os.environ.get("KB_N", 8820))
solver = os.environ.get("KB_SOLVER", bench.SOLVER)
which = os.environ.get("KB_MODEL", "clipper")
dev = torch.device("cuda", 0)
if which == "clipper":
    m = ex.diodeclipper(); P = [bench.sweep_params(B, 0, B)]
    if os.environ.get("KB_STIFF"):  # every warp holds the stiff corner of the sweep (k 224..255, j 0..7)
        idx = np.arange(B); P = [bench.sweep_params(65536, 0, 65536)[:, (224 + (idx % 256) % 32) + 256 * ((idx // 256) % 8)]]
    nu = 1
elif which == "birdie":
    m = ex.birdie(vol=0.8); P = None; nu = 1
elif which == "sallenkey":
    m = ex.sallenkey(fs=96000); P = None; nu = 1
r = BatchRunner(m, B, params=P, solver=solver)
row = torch.from_numpy(np.sin(2*np.pi*1000/44100*np.arange(N))).to(dev)
if which == "birdie":
    g = torch.Generator(device=dev); g.manual_seed(1)
    U = (0.2*torch.randn((B, N, 1), generator=g, device=dev, dtype=torch.float64)).clamp_(-1, 1)
else:
    U = row.reshape(1, N, 1).expand(B, N, 1).contiguous()
Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
for _ in range(2): r.run(U, Y, check_status=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 3
e0.record()
for _ in range(K): r.run(U, Y, check_status=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/K
st = r.stats()
print(json.dumps({"lib": os.path.basename(os.environ.get("ACMEB200_LIB", "default")), "model": which, "kernel": r.kernel_name[:30], "ms": round(ms, 3),
                  "Gsamples_s": round(B*N/ms/1e6, 2), "mean_iters": round(st["newton_iters"]/max(st["solves"],1), 3), "chk": float(Y[123, -1, 0])}))
