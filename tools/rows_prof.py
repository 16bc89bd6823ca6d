"""one launch of the rows kernel for ncu: superover, B and N from the environment"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
B = int(os.environ.get("KB_B", 1024)); N = int(os.environ.get("KB_N", 300)); kernel = os.environ.get("KB_KERNEL", "rows")
HC = "HomotopySolver{CachingSolver{SimpleSolver}}"
dev = torch.device("cuda", 0)
u = np.zeros((B, N, 4))
u[:, :, 0] = np.sin(2 * np.pi * 1000 / 44100 * np.arange(N))[None, :]
u[:, :, 1] = ((np.arange(B) % 128) + 0.5)[:, None] / 128
u[:, :, 2] = ((np.arange(B) // 128 % 64) + 0.5)[:, None] / 64
u[:, :, 3] = 1.0
U = torch.from_numpy(u).to(dev)
Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
r = BatchRunner(ex.superover(), B, solver=HC, kernel=kernel)
for _ in range(3):
    r.run(U, Y, check_status=False)
torch.cuda.synchronize()
print(r.kernel_name, r.stats()["newton_iters"] / r.stats()["solves"])
