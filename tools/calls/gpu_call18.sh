#!/bin/bash
# Round 2, call 18: kb-scale reads with scanfrag_kernel + gather_kernel: parity, timing, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_giant_reads.py tests/test_device_pack.py -m gpu -x -q > gpurun_out/c18_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c18_pytest.log
tail -15 gpurun_out/c18_pytest.log
run() { echo "== $*" | tee -a gpurun_out/c18_kbench.log; env "$@" timeout 300 python tools/kbench.py $CFG $N 4 2>&1 | tail -1 | cut -c1-200 | tee -a gpurun_out/c18_kbench.log; }
CFG=4 N=400000
run CLS_NO_FRAG=1
run CLS_X=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/c18_launches_cfg4.csv python tools/kbench.py 4 100000 1 > gpurun_out/c18_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/c18_launches_cfg4.csv') if l.startswith('"'))]
h=rows[0]; iK=h.index('Kernel Name'); iV=h.index('Metric Value'); iG=h.index('Grid Size')
for r in rows[1:22]:
    print(r[iK][:60].replace('void cls::',''), r[iG], r[iV])
PY
