"""how the learning solution cache grows over successive seconds of config 4 and what it does to the throughput"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
import bench
B = int(os.environ.get("KB_B", 1024)); N = 44100; secs = int(os.environ.get("KB_SECS", 4))
dev = torch.device("cuda", 0)
U = torch.from_numpy(bench.c4_inputs_np(0, B, N)).to(dev); Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
r = BatchRunner(ex.superover(), B, solver=bench.SOLVER)
prev = dict(newton_iters=0, solves=0, homotopy_solves=0)
for s in range(secs):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.run(U, Y, check_status=False); e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1)
    st = r.stats(); n, cap = r.cache_sizes()
    print(json.dumps(dict(second=s + 1, Msamples_s=round(B * N / ms / 1e3, 2), iters=round((st["newton_iters"] - prev["newton_iters"]) / (st["solves"] - prev["solves"]), 3),
                          homotopy=st["homotopy_solves"] - prev["homotopy_solves"], cache_mean=float(n.mean()), cache_max=int(n.max()), cache_min=int(n.min()), cap=cap)), flush=True)
    prev = st
