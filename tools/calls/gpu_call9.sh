#!/bin/bash
# Round 2, call 9: in-register merge at hand-over + grouped descent; bench N=1 config 3; ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan2.py tests/test_gpu_parity.py tests/test_routed.py tests/test_multi_device.py tests/test_zbuild_device.py -m gpu -x -q > gpurun_out/c9_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c9_pytest.log
tail -4 gpurun_out/c9_pytest.log
for CFG in 2 3; do
  timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c9_kbench.log
  CLS_DESCEND_GROUPS=0 timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c9_kbench.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan2_kernel|descend" -c 2 -o gpurun_out/c9_prof_cfg2 python tools/kbench.py 2 1000000 1 > gpurun_out/c9_ncu_cfg2.log 2>&1
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/c9_bench_cfg3.json 2> gpurun_out/c9_bench_cfg3.err; tail -c 600 gpurun_out/c9_bench_cfg3.json; tail -3 gpurun_out/c9_bench_cfg3.err
