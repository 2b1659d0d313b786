#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
usage: hotspots.py src_cs.csv [top_n]"""
import csv
import sys

def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out, hdr, cur_file = [], None, None
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        iS, iI = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    elif hdr and r and r[0].isdigit():
        out.append((cur_file, int(r[0]), r[1], num(r[iS]), num(r[iI])))
ts, ti = sum(o[3] for o in out) or 1, sum(o[4] for o in out) or 1
print(f"total stall samples {ts}, warp instructions {ti}")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{o[0]}:{o[1]:<4d} {100 * o[3] / ts:5.1f}% samples {100 * o[4] / ti:5.1f}% inst | {o[2].strip()[:100]}")
