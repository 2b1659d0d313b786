#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by named line ranges of
kernels.cu.  usage: regions.py src_cs.csv n_reads name:first-last [name:first-last ...]"""
import csv
import sys


def num(x):
    try:
        return int(float(x))
    except ValueError:
        return 0


rows = list(csv.reader(open(sys.argv[1])))
n_reads = float(sys.argv[2])
regions = []
for a in sys.argv[3:]:
    name, rng = a.split(":")
    lo, hi = rng.split("-")
    regions.append((name, int(lo), int(hi)))
out, hdr, cur = [], None, None
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        iS, iI = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    elif hdr and r and r[0].isdigit():
        out.append((cur, int(r[0]), num(r[iS]), num(r[iI])))
ts, ti = sum(o[2] for o in out) or 1, sum(o[3] for o in out) or 1
agg = {}
for f, l, s, i in out:
    key = f
    if f == "kernels.cu":
        key = "kernels.cu:other"
        for n, a, b in regions:
            if a <= l <= b:
                key = n
                break
    a = agg.setdefault(key, [0, 0])
    a[0] += s
    a[1] += i
print(f"total stall samples {ts}, warp instructions {ti} = {ti / n_reads:.0f} per read")
for k, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:36s} {100 * s / ts:5.1f}% samples {100 * i / ti:5.1f}% inst {i / n_reads:8.1f} inst/read")
