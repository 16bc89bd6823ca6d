"""summaries of an .ncu-rep: key raw metrics, stall mix, and executed instructions by region/opcode
usage: python tools/ncu_summary.py report.ncu-rep [nsamples_for_per_sample_counts]"""
import csv, subprocess, sys, io, collections, re
rep = sys.argv[1]; nsamp = float(sys.argv[2]) if len(sys.argv) > 2 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2] if len(rows) > 2 else rows[1]))
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.per_cycle_active",
        "smsp__issue_active.avg.per_cycle_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__icc_request_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct"]
for k in want:
    if k in d: print(f"{k:75s} {d[k]}")
print("--- stalls per issued instruction")
st = {k: float(v) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v}
for k, v in sorted(st.items(), key=lambda x: -x[1])[:10]:
    print(f"  {k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):25s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ia = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); isrc = hdr.index("Source")
data = [(r[isrc].strip(), int(r[ia]), int(r[isamp])) for r in rows[2:] if len(r) > isamp]
tot = sum(x[1] for x in data); ts = sum(x[2] for x in data)
print(f"--- executed warp instructions {tot} ({tot / nsamp if nsamp else 0:.0f} per sample), static {len(data)}")
ops = collections.Counter(); smp = collections.Counter()
for s, e, sa in data:
    m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_]+)', s)
    op = m.group(2) if m else "?"
    ops[op] += e; smp[op] += sa
for op, e in ops.most_common(22):
    print(f"  {op:12s} exec {e / tot * 100:5.1f}%  stall-samples {smp[op] / ts * 100:5.1f}%" + (f"  per sample {e / nsamp:8.1f}" if nsamp else ""))
