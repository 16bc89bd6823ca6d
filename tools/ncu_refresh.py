"""Refreshes the ncu-derived roofline inputs that bench.py reports (profiles/r2/traffic_*.json): DRAM bytes and executed
FP64 flops per circuit-sample of the dominant kernel of configs 2, 3 and 5, each from ONE profiled launch of a shortened
run, stamped with the content hash of csrc/ (bench.src_stamp) so that bench.py drops the figures the moment the kernels
change.  Run on the GPU box:  python tools/ncu_refresh.py   (writes gpurun_out/ too, for the trip home)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum",
           "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
           "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.per_cycle_active", "sm__warps_active.avg.per_cycle_active"]


def capture(env, script, kernel_regex, skip, samples_per_launch):
    cmd = ["ncu", "--metrics", ",".join(METRICS), "--clock-control", "none", "-k", f"regex:{kernel_regex}", "-s", str(skip), "-c", "1",
           "--csv", sys.executable, os.path.join(ROOT, script)]
    res = subprocess.run(cmd, env=dict(os.environ, **env), capture_output=True, text=True, timeout=900)
    lines = [l for l in res.stdout.splitlines() if l.startswith('"')]
    rows = list(csv.reader(io.StringIO("\n".join(lines))))
    if len(rows) < 2:
        raise RuntimeError("no ncu rows:\n" + res.stdout[-2000:] + res.stderr[-2000:])
    hdr = rows[0]; iname, ival, iunit = hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    vals, units = {}, {}
    for r in rows[1:]:
        vals[r[iname]] = float(r[ival].replace(",", "")); units[r[iname]] = r[iunit]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    dram = sum(vals[k] * scale.get(units[k], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    flops = 2 * vals["smsp__sass_thread_inst_executed_op_dfma_pred_on.sum"] + vals["smsp__sass_thread_inst_executed_op_dmul_pred_on.sum"] + \
        vals["smsp__sass_thread_inst_executed_op_dadd_pred_on.sum"]
    return {"kernel": rows[1][hdr.index("Kernel Name")][:80], "capture": "ncu --metrics (one launch, %d circuit-samples) of %s %s" % (
                samples_per_launch, script, " ".join(f"{k}={v}" for k, v in env.items())),
            "src_stamp": bench.src_stamp(), "samples_captured": samples_per_launch, "dram_bytes_captured_launch": dram,
            "dram_bytes_per_sample": dram / samples_per_launch, "ratio_to_algorithmic": dram / (16.0 * samples_per_launch),
            "fp64_flops_per_sample": flops / samples_per_launch, "warp_instructions_per_sample": vals["smsp__inst_executed.sum"] / samples_per_launch * 32,
            "fp64_pipe_pct": vals["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"],
            "issue_active": vals["smsp__issue_active.avg.per_cycle_active"], "warps_active_per_sm": vals["sm__warps_active.avg.per_cycle_active"],
            "ncu_kernel_ms": vals["gpu__time_duration.sum"] * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(units["gpu__time_duration.sum"], 1e-6)}


out = {}
jobs = [("traffic_clipper.json", dict(KB_MODEL="clipper", KB_N="4410"), "tools/kbench_one.py", "k_tpi", 4, 65536 * 4410),
        ("traffic_linear.json", dict(KB_MODEL="sallenkey", KB_N="9600"), "tools/kbench_one.py", "k_tpi", 4, 65536 * 9600),
        ("traffic_birdie.json", dict(KB_WARM="44100", KB_N="2205"), "tools/birdie_prof.py", "k_tpi", 2, 32768 * 2205)]
for name, env, script, rx, skip, nsamp in jobs:
    try:
        c = capture(env, script, rx, skip, nsamp)
    except Exception as e:
        print(name, "FAILED", e); continue
    if name == "traffic_linear.json":
        c["dram_bytes_per_launch"] = c["dram_bytes_per_sample"] * 65536 * 96000
    if name == "traffic_birdie.json":
        c["dram_bytes_per_launch"] = c["dram_bytes_per_sample"] * 32768 * 44100
    for d in (os.path.join(ROOT, "profiles", "r2"), os.path.join(ROOT, "gpurun_out")):
        os.makedirs(d, exist_ok=True)
        json.dump(c, open(os.path.join(d, name), "w"), indent=1)
    print(name, json.dumps({k: c[k] for k in ("dram_bytes_per_sample", "ratio_to_algorithmic", "fp64_flops_per_sample", "warp_instructions_per_sample", "fp64_pipe_pct")}))
