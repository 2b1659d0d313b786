#!/bin/bash
# Round 2, call 4: A/B of the scan2 switches (parity of the default build first), L2 fetch granularity, config 2 and 3.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan2.py tests/test_gpu_parity.py tests/test_routed.py -m gpu -x -q > gpurun_out/c4_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c4_pytest.log
tail -4 gpurun_out/c4_pytest.log
for CFG in 2 3; do
  timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c4_kbench.log
  for V in nodirect olddecode noprefetch block4 nopred nobloom ld128; do
    CLASSEQ_B200_LIB=$PWD/classeq2_b200/libclasseq_b200_$V.so timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c4_kbench.log
  done
  for G in 32 128; do
    echo "CLS_L2_FETCH=$G" | tee -a gpurun_out/c4_kbench.log
    CLS_L2_FETCH=$G timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c4_kbench.log
    CLS_L2_FETCH=$G CLASSEQ_B200_LIB=$PWD/classeq2_b200/libclasseq_b200_ld128.so timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c4_kbench.log
  done
done
