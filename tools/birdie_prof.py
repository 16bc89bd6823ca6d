"""config 5 in small: birdie(vol=0.8), white noise, learning cache warmed by `KB_WARM` samples, then one launch for ncu"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
import bench
B = int(os.environ.get("KB_B", 32768)); N = int(os.environ.get("KB_N", 1102)); W = int(os.environ.get("KB_WARM", 22050))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0xACE5EED)
r = BatchRunner(ex.birdie(vol=0.8), B, solver=bench.SOLVER)
U = (0.2 * torch.randn((B, W, 1), generator=g, device=dev, dtype=torch.float64)).clamp_(-1, 1)
Y = torch.empty((B, W, 1), dtype=torch.float64, device=dev)
r.run(U, Y, check_status=False); torch.cuda.synchronize()
U2 = U[:, :N].contiguous(); Y2 = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.run(U2, Y2, check_status=False); e1.record(); torch.cuda.synchronize()
print(r.kernel_name, "Gsamples/s", B * N / e0.elapsed_time(e1) / 1e6, "stored", r.cache_sizes()[0].mean())
