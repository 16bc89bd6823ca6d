import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
B = int(os.environ.get("KB_B", 8192)); N = int(os.environ.get("KB_N", 2000)); kernel = os.environ.get("KB_KERNEL", "auto")
dev = torch.device("cuda", 0)
m = ex.superover()
r = BatchRunner(m, B, kernel=kernel)
U = torch.zeros((B, N, 4), dtype=torch.float64, device=dev)
U[:, :, 0] = torch.sin(2*np.pi*1000/44100*torch.arange(N, device=dev, dtype=torch.float64))[None, :]
k = torch.arange(B, device=dev)
U[:, :, 1] = (((k % 128) + 0.5) / 128)[:, None]
U[:, :, 2] = (((k // 128) % 64 + 0.5) / 64)[:, None]
U[:, :, 3] = 1.0
Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
r.run(U, Y, check_status=False); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 2
e0.record()
for _ in range(K): r.run(U, Y, check_status=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/K
st = r.stats()
print(json.dumps({"model": "superover", "B": B, "N": N, "kernel": r.kernel_name[:40], "ms": round(ms, 2), "Msamples_s": round(B*N/ms/1e3, 2),
                  "mean_iters": round(st["newton_iters"]/max(st["solves"],1), 3), "homotopy": st["homotopy_solves"], "bad": int((r.status()[0]!=0).sum())}))
