"""config 3 (Sallen-Key, per-instance matrices, B=65536, 96 kHz) and config 2 (clipper) kernel timing for
one library build (ACMEB200_LIB); N shortened (the kernels are time loops).  KB_LAYOUT=sample times the same
work on sample-major streams (ACMEB200_SAMPLE_MAJOR); the checksums must not change."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
import bench
dev = torch.device("cuda", 0)
LAYOUT = os.environ.get("KB_LAYOUT", "instance")
def streams(row, N):  # device input/output streams of the chosen layout, and the accessor of (instance, sample)
    if LAYOUT == "sample":
        U = row.reshape(N, 1, 1).expand(N, B, 1).contiguous()
        return U, torch.empty_like(U), (lambda Y, b, n: float(Y[n, b, 0]))
    U = row.reshape(1, N, 1).expand(B, N, 1).contiguous()
    return U, torch.empty_like(U), (lambda Y, b, n: float(Y[b, n, 0]))
def timed(r, U, Y, steps=3, warm=2):
    for _ in range(warm): r.run(U, Y, check_status=False, layout=LAYOUT)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): r.run(U, Y, check_status=False, layout=LAYOUT)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
B = 65536
N = int(os.environ.get("KB_N", 19200))
base = ex.sallenkey(fs=96000)
mats = {k: [] for k in ("a", "b", "x0", "dy", "ey", "y0")}
for k in range(16):
    R = 10 ** (3 + 2 * (k % 4) / 3); kap = 10 ** (1.3 * (k // 4) / 3)
    mk = ex.sallenkey(fs=96000, r1=R, r2=R, c1=10e-9 * kap, c2=10e-9 / kap)
    for key in mats: mats[key].append(getattr(mk, key))
ov = {k: np.tile(np.stack(v, axis=-1), (1,) * (np.ndim(v[0])) + (B // 16,)) for k, v in mats.items()}
r = BatchRunner(base, B, overrides=ov)
U, Y, at = streams(torch.from_numpy(np.sin(2 * np.pi * 1000 / 96000 * np.arange(N))).to(dev), N)
ms = timed(r, U, Y)
out = {"lib": os.path.basename(os.environ.get("ACMEB200_LIB", "default")), "layout": LAYOUT, "cfg3_Gs": round(B * N / ms / 1e6, 1), "cfg3_TBs": round(16 * B * N / ms / 1e9, 3), "chk3": at(Y, 777, N - 1)}
del U, Y; r.close(); torch.cuda.empty_cache()
N = int(os.environ.get("KB_N2", 8820))
m = ex.diodeclipper(); P = bench.sweep_params(B, 0, B)
r = BatchRunner(m, B, params=[P])
U, Y, at = streams(torch.from_numpy(np.sin(2 * np.pi * 1000 / 44100 * np.arange(N))).to(dev), N)
ms = timed(r, U, Y)
out.update({"cfg2_Gs": round(B * N / ms / 1e6, 2), "chk2": at(Y, 777, N - 1)})
print(json.dumps(out))
