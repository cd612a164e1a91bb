#!/usr/bin/env python
"""Development helper: per-section (gather / sweeps) instruction and stall-sample shares of an ncu report.
usage: scripts/ncu_sections.py gpurun_out/prof.ncu-rep"""
import csv, io, subprocess, sys, re
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
files = {}
cur = None
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        cur = r[1]; files[cur] = []; continue
    if r[0] == "Line No": hdr = r; continue
    if cur and r[0].isdigit() and hdr:
        iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
        try: files[cur].append((int(r[0]), r[1], int(r[iS]), int(r[iI])))
        except Exception: pass
totI = sum(x[3] for f in files.values() for x in f); totS = sum(x[2] for f in files.values() for x in f)
print("total instr %.3e samples %d" % (totI, totS))
main = [f for f in files if f.endswith("kgw_device.cuh")][0]
text = open(main).read().split("\n")
marks = []
for i, l in enumerate(text, 1):
    m = re.search(r"// -{10,} (.*)", l)
    if m: marks.append((i, m.group(1)))
marks = [(1, "helpers")] + marks + [(10**9, "end")]
for (a, name), (b, _) in zip(marks, marks[1:]):
    I = sum(x[3] for x in files[main] if a <= x[0] < b); S = sum(x[2] for x in files[main] if a <= x[0] < b)
    print(f"{name[:50]:50s} lines {a:4d}-{min(b,len(text)):4d}  I={100*I/totI:5.1f}%  S={100*S/totS:5.1f}%")
for f in files:
    if f == main: continue
    I = sum(x[3] for x in files[f]); S = sum(x[2] for x in files[f])
    print(f"{f[-50:]:50s}  I={100*I/totI:5.1f}%  S={100*S/totS:5.1f}%")
if len(sys.argv) > 3:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    unit = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
    for f in files:
        for ln, s, S, I in files[f]:
            if (f == main and a <= ln < b and I > 0) or (f != main and a == 0 and I > 0):
                print(f"{f[-14:]:14s} {ln:4d} I/unit={I/unit:7.1f} S={100*S/totS:4.1f}%  {s[:110]}")
