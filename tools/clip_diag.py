import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
import bench
B, N = 65536, 8820
dev = torch.device("cuda", 0)
row = torch.from_numpy(np.sin(2*np.pi*1000/44100*np.arange(N))).to(dev)
U = row.reshape(1, N, 1).expand(B, N, 1).contiguous(); Y = torch.empty_like(U)
for solver in ("HomotopySolver{SimpleSolver}", bench.SOLVER):
    r = BatchRunner(ex.diodeclipper(), B, params=[bench.sweep_params(B, 0, B)], solver=solver)
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r.run(U, Y, check_status=False); e1.record(); torch.cuda.synchronize()
        info = r.cache_info()
        print(solver[:24], "launch", rep, "Gs/s %.2f" % (B*N/e0.elapsed_time(e1)/1e6), "num_ps hist", np.bincount(info["num_ps"])[:12].tolist(), "newc>0:", int((info["new_count"]>0).sum()),
              "treen hist", np.bincount(info["tree_n"])[:8].tolist())
    r.close()
