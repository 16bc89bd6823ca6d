"""Where does the thread-per-instance kernel's tail come from?  Needs a library built with ACME_TPI_PROF
(tools/build_variants.py prof:64:8:8:2:2:ACME_TPI_PROF=1; ACMEB200_LIB=tools/libs/lib_prof.so).
Per-warp cycle counts and event counters of one steady-state launch of config 2 (or KB_MODEL=birdie)."""
import os, sys, json, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from acme_jl_b200 import BatchRunner, examples as ex
from acme_jl_b200._lib import lib
import bench
B = int(os.environ.get("KB_B", 65536)); N = int(os.environ.get("KB_N", 8820))
which = os.environ.get("KB_MODEL", "clipper")
dev = torch.device("cuda", 0)
L = lib()
has_prof = hasattr(L, "acmeb200_diag_tpi_prof")
def fetch(nw):
    out = np.zeros((nw, 8), dtype=np.uint64)
    L.acmeb200_diag_tpi_prof(out.ctypes.data_as(ctypes.c_void_p), nw)
    return out
def experiment(tag, solver, P):
    if which == "clipper":
        r = BatchRunner(ex.diodeclipper(), B, params=[P], solver=solver)
        row = torch.from_numpy(np.sin(2*np.pi*1000/44100*np.arange(N))).to(dev)
        U = row.reshape(1, N, 1).expand(B, N, 1).contiguous()
    else:
        r = BatchRunner(ex.birdie(vol=0.8), B, solver=solver)
        g = torch.Generator(device=dev); g.manual_seed(1)
        U = (0.2*torch.randn((B, N, 1), generator=g, device=dev, dtype=torch.float64)).clamp_(-1, 1)
    Y = torch.empty((B, N, 1), dtype=torch.float64, device=dev)
    for _ in range(int(os.environ.get("KB_WARM", 5))): r.run(U, Y, check_status=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r.run(U, Y, check_status=False); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rec = {"tag": tag, "solver": solver[:24], "ms": round(ms, 2), "Gs": round(B*N/ms/1e6, 2)}
    if has_prof:
        nw = min(B // 32, 8192)
        p = fetch(nw).astype(np.int64)
        cyc = p[:, 0]
        q = np.percentile(cyc, [0, 50, 90, 99, 100]) / N
        rec["cyc_per_sample_min_med_p90_p99_max"] = [round(float(x), 1) for x in q]
        rec["cold_cyc_share_total"] = round(float(p[:, 1].sum() / cyc.sum()), 4)
        order = np.argsort(-cyc)[:12]
        rows = []
        for w in order:
            rows.append({"warp": int(w), "k0": int((w*32) % 256), "j": int((w*32)//256), "cyc/s": round(float(cyc[w]/N), 1), "cold_cyc/s": round(float(p[w, 1]/N), 1),
                         "reo": int(p[w, 2] >> 32), "tie": int(p[w, 2] & 0xffffffff), "scan": int(p[w, 3] >> 32), "small": int(p[w, 3] & 0xffffffff),
                         "nontriv": int(p[w, 4] >> 32), "nontriv_max_lane": int(p[w, 4] & 0xffffffff), "cold": int(p[w, 5] >> 32), "reb": int(p[w, 5] & 0xffffffff),
                         "stores": int(p[w, 6]), "sm": int(p[w, 7])})
        rec["slowest"] = rows
        # slow-warp census: warps above 1.15 x median
        slow = cyc > 1.15 * np.median(cyc)
        rec["n_slow_warps"] = int(slow.sum()); rec["slow_have_nontriv"] = int(((p[:, 4] >> 32)[slow] > 0).sum())
        rec["warps_with_nontriv"] = int(((p[:, 4] >> 32) > 0).sum())
        # SM finish-time spread: per SM the max warp cycles
        sms = p[:, 7]
        per_sm = np.array([cyc[sms == s].max() for s in np.unique(sms)])
        rec["per_sm_maxcyc_min_med_max"] = [round(float(x)/N, 1) for x in (per_sm.min(), np.median(per_sm), per_sm.max())]
        info = r.cache_info()
        rec["num_ps_hist"] = np.bincount(info["num_ps"])[:10].tolist(); rec["treen_hist"] = np.bincount(info["tree_n"])[:10].tolist()
        rec["newc_pos"] = int((info["new_count"] > 0).sum())
    print(json.dumps(rec), flush=True)
    r.close()
HC, H = bench.SOLVER, "HomotopySolver{SimpleSolver}"
if which == "clipper":
    P = bench.sweep_params(B, 0, B)
    experiment("full sweep", HC, P)
    experiment("full sweep", H, P)
    idx = np.arange(B); k, j = idx % 256, idx // 256
    easy = (np.minimum(k, 200) + 256 * np.maximum(j, 40))
    experiment("stiff corner replaced (k<=200, j>=40)", HC, P[:, easy])
    stiff = (224 + k % 32) + 256 * (j % 8)
    experiment("all stiff (k 224..255, j 0..7)", HC, P[:, stiff])
    experiment("all stiff (k 224..255, j 0..7)", H, P[:, stiff])
else:
    experiment("birdie", HC, None)
