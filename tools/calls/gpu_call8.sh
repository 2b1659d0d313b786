#!/bin/bash
# Round 2, call 8: grouped descent (several reads per warp) parity + A/B; full GPU suite; ncu of the final kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan2.py tests/test_gpu_parity.py tests/test_routed.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/c8_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c8_pytest.log
tail -6 gpurun_out/c8_pytest.log
for CFG in 2 3; do
  timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c8_kbench.log
  CLS_DESCEND_GROUPS=0 timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c8_kbench.log
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c8_pytest_all.log 2>&1; echo "rc=$?" >> gpurun_out/c8_pytest_all.log
tail -4 gpurun_out/c8_pytest_all.log
for CFG in 2 3; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan2_kernel|descend" -c 2 -o gpurun_out/c8_prof_cfg$CFG python tools/kbench.py $CFG 1000000 1 > gpurun_out/c8_ncu_cfg$CFG.log 2>&1
done
ls -la gpurun_out/c8_prof*
