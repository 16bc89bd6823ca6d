"""stall samples / executed instructions per CUDA source line of an .ncu-rep (needs -lineinfo and --import-source on)
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, io, collections, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file = cur_line = hdr = None; src = {}
stat = collections.Counter(); exe = collections.Counter(); samp = collections.Counter()
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur_file = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; iex = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr is None or len(r) <= iex: continue
    if r[0] != "": cur_line = int(r[0]); src[(cur_file, cur_line)] = r[1].strip()[:100]; continue
    k = (cur_file, cur_line); stat[k] += 1
    try: exe[k] += int(r[iex]); samp[k] += int(r[isamp])
    except ValueError: pass
ts = sum(samp.values()) or 1; te = sum(exe.values()) or 1
print("static SASS", sum(stat.values()), "executed", te, "samples", ts)
for k, s in samp.most_common(top):
    print(f"{s / ts * 100:5.1f}% smp {exe[k] / te * 100:5.1f}% exe {stat[k]:5d} sass  {k[0]}:{k[1]}  {src.get(k, '')}")
