#!/bin/bash
# Round 2, call 22: adaptive hand-out at the end of the descent launch: parity, end-to-end and resident timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan2.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/c22_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c22_pytest.log
tail -3 gpurun_out/c22_pytest.log
for i in 1 2; do PIN=1 timeout 300 python tools/e2e_bench.py 10000000 3 2>&1 | tail -1 | cut -c1-260 | tee -a gpurun_out/c22_e2e.log; done
PIN=1 timeout 300 python tools/e2e_bench.py 1250000 3 2>&1 | tail -1 | cut -c1-260 | tee -a gpurun_out/c22_e2e.log
timeout 200 python tools/kbench.py 3 4000000 6 2>&1 | tail -1 | cut -c1-110 | tee -a gpurun_out/c22_kbench.log
timeout 200 python tools/kbench.py 2 1000000 10 2>&1 | tail -1 | cut -c1-110 | tee -a gpurun_out/c22_kbench.log
