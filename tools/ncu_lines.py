#!/usr/bin/env python
"""Rank CUDA source lines of a kernel by executed instructions / stall samples from an
`ncu --page source --print-source cuda,sass --csv` export.  usage: ncu_lines.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
hdr = None
data = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        sm = hdr.index("# Samples")
        continue
    if hdr is None or r[0] == "" or r[0] == "Function Name":
        continue
    try:
        data.append((int(r[ie]), int(r[sm]), cur_file, r[0], r[1].strip()[:100]))
    except (ValueError, IndexError):
        pass
tot = sum(d[0] for d in data)
tots = sum(d[1] for d in data)
print("total warp-instructions", tot, "samples", tots)
for d in sorted(data, key=lambda d: -d[0])[:top]:
    print(f"{100 * d[0] / tot:5.1f}% inst {100 * d[1] / max(tots, 1):5.1f}% smp  {d[2]}:{d[3]:>4}  {d[4]}")
