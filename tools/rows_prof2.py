"""config 4 for ncu: superover, KB_B instances, the solution store learnt over three launches of KB_N samples, then one
short launch of KB_PROF_N samples (the one to capture: -k regex:k_rows -s 3 -c 1)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
import bench
B = int(os.environ.get("KB_B", 1024)); N = int(os.environ.get("KB_N", 44100)); NP_ = int(os.environ.get("KB_PROF_N", 600))
dev = torch.device("cuda", 0)
U = torch.from_numpy(bench.c4_inputs_np(0, B, N)).to(dev)
Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
r = BatchRunner(ex.superover(), B, solver=bench.SOLVER)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.run(U, Y, check_status=False); e1.record(); torch.cuda.synchronize()
    print("warm", B * N / e0.elapsed_time(e1) / 1e3, "Msamples/s")
U2 = U[:, :NP_].contiguous(); Y2 = torch.empty((B, NP_, 1), dtype=torch.float64, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); r.run(U2, Y2, check_status=False); e1.record(); torch.cuda.synchronize()
info = r.cache_info()
print(r.kernel_name, "Msamples/s", B * NP_ / e0.elapsed_time(e1) / 1e3, "stored mean", info["num_ps"].mean(), "tree_n mean", info["tree_n"].mean(), "flags", np.bincount(info["flags"]))
