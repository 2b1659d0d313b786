#!/bin/bash
# Round 2, call 5: scan2 after the uniform-descriptor fix; parity, A/B, ncu on config 2 and config 3; multi-device C ABI test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan2.py tests/test_gpu_parity.py tests/test_routed.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/c5_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c5_pytest.log
tail -6 gpurun_out/c5_pytest.log
for CFG in 2 3; do
  timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c5_kbench.log
  for V in olddecode noprefetch nobloom m4 nodirect; do
    CLASSEQ_B200_LIB=$PWD/classeq2_b200/libclasseq_b200_$V.so timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c5_kbench.log
  done
done
for CFG in 2 3; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan2_kernel|descend_kernel" -c 2 -o gpurun_out/c5_prof_cfg$CFG python tools/kbench.py $CFG 1000000 1 > gpurun_out/c5_ncu_cfg$CFG.log 2>&1
tail -1 gpurun_out/c5_ncu_cfg$CFG.log
done
