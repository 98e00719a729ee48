#!/usr/bin/env python
"""Aggregate an ncu report's source page by CUDA source line: executed warp instructions and stall samples.
    python benchmarks/ncu_source_lines.py file.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(raw))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
cols = rows[hdr]
ix_inst, ix_samp = cols.index("Instructions Executed"), cols.index("# Samples")
lines, cur = [], None
for r in rows[hdr + 1:]:
    if len(r) <= ix_inst:
        continue
    if r[0].isdigit():                         # a CUDA source line (aggregated over its SASS)
        num = lambda v: int(v) if v.isdigit() else 0
        lines.append((int(r[0]), r[1].strip(), num(r[ix_inst]), num(r[ix_samp])))
tot_i, tot_s = sum(l[2] for l in lines), sum(l[3] for l in lines)
print(f"total warp instructions {tot_i}, samples {tot_s}")
for ln, src, ins, smp in sorted(lines, key=lambda l: -l[2])[:top]:
    print(f"{100 * ins / max(tot_i, 1):5.1f}% inst {100 * smp / max(tot_s, 1):5.1f}% samp  L{ln:<4d} {src[:110]}")
