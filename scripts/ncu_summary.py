#!/usr/bin/env python
"""Development helper: summarise an ncu report (raw metrics + per-source-line hot spots).
usage: scripts/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread ', 'smsp__inst_executed.sum ', 'smsp__issue_active.avg.pct', 'sm__inst_executed_pipe_fp64.sum.pct',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct', 'smsp__average_warp_latency_per_inst_issued', 'launch__grid_size',
        'per_issue_active.ratio', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.sum.pct',
        'sm__inst_executed_pipe_fma.sum.pct', 'launch__occupancy_limit', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum ']
for h, u, v in zip(hdr, units, vals):
    if any(k in h + ' ' for k in keys):
        print(f"{h:90s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[2]
iS, iI = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall = {n: hdr.index(n) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
lines = []
for r in rows[3:]:
    if r and r[0].isdigit():
        try:
            lines.append((int(r[0]), r[1], int(r[iS]), int(r[iI]), {k: int(r[v] or 0) for k, v in stall.items()}))
        except Exception:
            pass
totS, totI = sum(x[2] for x in lines), sum(x[3] for x in lines)
print(f"total samples {totS}  total warp instructions {totI}")
print("-- by instructions")
for ln, s, sm, i, st in sorted(lines, key=lambda x: -x[3])[:nl]:
    print(f"{ln:4d} I={100*i/totI:5.1f}% S={100*sm/totS:5.1f}%  {s[:105]}")
print("-- by stall samples")
for ln, s, sm, i, st in sorted(lines, key=lambda x: -x[2])[:nl]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{ln:4d} S={100*sm/totS:5.1f}% I={100*i/totI:5.1f}% {[(k[6:], v) for k, v in top]}  {s[:80]}")
