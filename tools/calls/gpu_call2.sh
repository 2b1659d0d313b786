#!/bin/bash
# Round 2, call 2: parity of scan2_kernel (default build: 5 CTAs/SM, and the 4-CTA variant), timing, ncu capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan2.py tests/test_gpu_parity.py tests/test_routed.py -m gpu -x -q > gpurun_out/c2_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c2_pytest.log
tail -15 gpurun_out/c2_pytest.log
for CFG in 2 3; do
  N=1000000
  timeout 200 python tools/kbench.py $CFG $N 10 2>&1 | tail -1 | tee -a gpurun_out/c2_kbench.log
  for V in m4; do
    CLASSEQ_B200_LIB=$PWD/classeq2_b200/libclasseq_b200_$V.so timeout 200 python tools/kbench.py $CFG $N 10 2>&1 | tail -1 | tee -a gpurun_out/c2_kbench.log
  done
done
CLASSEQ_B200_LIB=$PWD/classeq2_b200/libclasseq_b200_m4.so timeout 300 python -m pytest tests/test_scan2.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/c2_pytest_m4.log
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/c2_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/c2_pytest_all.log
tail -4 gpurun_out/c2_pytest_all.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan2_kernel|descend_kernel" -c 2 -o gpurun_out/c2_prof python tools/kbench.py 2 1000000 1 > gpurun_out/c2_ncu.log 2>&1
tail -3 gpurun_out/c2_ncu.log
ls -la gpurun_out/*.ncu-rep
