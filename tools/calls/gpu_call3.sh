#!/bin/bash
# Round 2, call 3: scan2 with block descriptor loads, fast decode, direct hand-over, chase loop; parity, timing, ncu (config 3).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan2.py tests/test_gpu_parity.py tests/test_routed.py -m gpu -x -q > gpurun_out/c3_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c3_pytest.log
tail -6 gpurun_out/c3_pytest.log
for CFG in 2 3; do
  timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | tee -a gpurun_out/c3_kbench.log
  CLASSEQ_B200_LIB=$PWD/classeq2_b200/libclasseq_b200_m4.so timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | tee -a gpurun_out/c3_kbench.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan2_kernel|descend_kernel" -c 2 -o gpurun_out/c3_prof python tools/kbench.py 3 1000000 1 > gpurun_out/c3_ncu.log 2>&1
tail -2 gpurun_out/c3_ncu.log
