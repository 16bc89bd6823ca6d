"""learning cache of the thread-per-instance kernel on birdie(vol=0.8) with white noise, second by second"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
import bench
B = int(os.environ.get("KB_B", 8192)); N = 44100; secs = int(os.environ.get("KB_SECS", 4))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0xACE5EED)
r = BatchRunner(ex.birdie(vol=0.8), B, solver=bench.SOLVER)
Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
prev = dict(newton_iters=0, solves=0, homotopy_solves=0)
for s in range(secs):
    U = (0.2 * torch.randn((B, N, 1), generator=g, device=dev, dtype=torch.float64)).clamp_(-1, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.run(U, Y, check_status=False); e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)
    st = r.stats(); n, cap = r.cache_sizes()
    print(json.dumps(dict(second=s + 1, Gsamples_s=round(B * N / ms / 1e6, 3), iters=round((st["newton_iters"] - prev["newton_iters"]) / (st["solves"] - prev["solves"]), 3),
                          homotopy=st["homotopy_solves"] - prev["homotopy_solves"], stored_mean=float(n.mean()), stored_max=int(n.max()), cap=cap)), flush=True)
    prev = st
