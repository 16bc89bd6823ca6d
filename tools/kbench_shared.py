import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
import bench
B, N = 65536, 8820
solver = os.environ.get("KB_SOLVER", bench.SOLVER)
dev = torch.device("cuda", 0)
m = ex.diodeclipper(); P = [bench.sweep_params(B, 0, B)]
r = BatchRunner(m, B, params=P, solver=solver)
U = torch.from_numpy(np.sin(2*np.pi*1000/44100*np.arange(N))).to(dev).reshape(N, 1).contiguous()
Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
for _ in range(2): r.run(U, Y, check_status=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): r.run(U, Y, check_status=False)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"shared_u": True, "solver": solver[:14], "ms": round(e0.elapsed_time(e1)/3, 3)}))
