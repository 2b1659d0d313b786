#!/bin/bash
# Round 2, call 16: which kernels run for config 4, and for how long (ncu launch list)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c16_launches_cfg4.csv python tools/kbench.py 4 100000 1 > gpurun_out/c16_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/c16_launches_cfg4.csv') if l.startswith('"'))]
h=rows[0]; iK=h.index('Kernel Name'); iV=h.index('Metric Value'); iG=h.index('Grid Size')
for r in rows[1:]:
    print(r[iK][:70].replace('void cls::',''), r[iG], r[iV])
PY
