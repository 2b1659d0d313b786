#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of the placement kernels (scan2_kernel + descend kernel of one step) into the
per-read counters bench.py reports as `roofline.secondary` / `roofline.traffic` (profiles/ncu_counters.json).

usage: summarize_ncu.py <capture.ncu-rep> <config> <n_reads> [--update]      (needs `ncu` on PATH; no GPU)
"""
import csv
import io
import json
import os
import subprocess
import sys

rep, config, n_reads = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
M = {"t_ms": "gpu__time_duration.sum", "inst": "smsp__inst_executed.sum", "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "alu": "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "fmaheavy": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
     "dram_r": "dram__bytes_read.sum", "dram_w": "dram__bytes_write.sum", "l2_sectors": "lts__t_sectors_srcunit_tex.sum",
     "l2_hit": "lts__t_sector_hit_rate.pct", "regs": "launch__registers_per_thread", "lanes": "smsp__thread_inst_executed_per_inst_executed.ratio",
     "warps": "smsp__warps_active.avg.per_cycle_active"}
units = dict(zip(hdr, rows[1]))


def val(d, key):
    v, u = float(d[M[key]]), units[M[key]]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
    return v * scale


kernels = []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    k = {"kernel": d["Kernel Name"].split("(")[0].replace("void ", "").replace("unnamed>::", "cls::"), "grid": d.get("Grid Size"),
         "time_ms": round(val(d, "t_ms"), 4), "warp_instructions_per_read": round(val(d, "inst") / n_reads, 1),
         "issue_active_pct": round(val(d, "issue"), 1), "alu_pipe_pct": round(val(d, "alu"), 1), "fmaheavy_pipe_pct": round(val(d, "fmaheavy"), 1),
         "active_lanes": round(val(d, "lanes"), 1), "warps_per_scheduler": round(val(d, "warps"), 1), "registers": int(val(d, "regs")),
         "dram_bytes_per_read": round((val(d, "dram_r") + val(d, "dram_w")) / n_reads, 1),
         "l2_bytes_per_read": round(val(d, "l2_sectors") * 32 / n_reads, 1), "l2_hit_pct": round(val(d, "l2_hit"), 1)}
    kernels.append(k)
tot_t = sum(k["time_ms"] for k in kernels)
entry = {"capture": os.path.relpath(rep, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "reads_in_capture": n_reads,
         "step_ms_in_capture": round(tot_t, 4),
         "warp_instructions_per_read": round(sum(k["warp_instructions_per_read"] for k in kernels), 1),
         "issue_active_pct": round(sum(k["issue_active_pct"] * k["time_ms"] for k in kernels) / tot_t, 1),
         "alu_pipe_pct": round(sum(k["alu_pipe_pct"] * k["time_ms"] for k in kernels) / tot_t, 1),
         "dram_bytes_per_read": round(sum(k["dram_bytes_per_read"] for k in kernels), 1),
         "l2_bytes_per_read": round(sum(k["l2_bytes_per_read"] for k in kernels), 1), "kernels": kernels}
dram_gbs = entry["dram_bytes_per_read"] * n_reads / (tot_t / 1e3) / 1e9
entry["dram_gbs_in_capture"] = round(dram_gbs, 1)
entry["limiter"] = ("HBM sectors (random 32-byte probes fetched at DRAM granularity) together with the ALU pipe" if dram_gbs > 3000
                    else "ALU pipe / instruction issue (integer hashing, compares, selects); the table is L2-resident")
print(json.dumps(entry, indent=1))
if "--update" in sys.argv:
    p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_counters.json")
    cur = json.load(open(p)) if os.path.exists(p) else {}
    cur[f"config{config}"] = entry
    json.dump(cur, open(p, "w"), indent=1)
    print("updated", p)
