"""top stalled SASS instructions of an .ncu-rep with their innermost source line and a few instructions of context
usage: python tools/ncu_sass_top.py report.ncu-rep [top_n] [context]"""
import csv, subprocess, io, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30; ctx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file = cur_line = hdr = None
seen = {}
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"): cur_file = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; iex = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); ia = hdr.index("Address"); continue
    if hdr is None or len(r) <= iex: continue
    if r[0] != "": cur_line = int(r[0]); continue
    try: seen.setdefault(r[ia], [int(r[isamp]), int(r[iex]), r[3].strip()[:70], []])[3].append(f"{cur_file}:{cur_line}")
    except ValueError: pass
addrs = sorted(seen, key=lambda a: int(a, 16))
tot = sum(v[0] for v in seen.values())
order = sorted(seen, key=lambda a: -seen[a][0])[:top]
for a in order:
    i = addrs.index(a)
    for j in range(max(0, i - ctx), i + 1):
        v = seen[addrs[j]]
        print(("  " if j < i else "> ") + f"{v[0] / tot * 100:5.2f}% {v[1]:>10d} {v[2]:70s} {' < '.join(v[3][:3])}")
    if ctx: print()
