#!/bin/bash
# Round 2, call 15: two-phase kb-scale path (scanfrag_kernel + prehashed placement kernel): parity, config 4 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_giant_reads.py tests/test_scan2.py -m gpu -x -q > gpurun_out/c15_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c15_pytest.log
tail -15 gpurun_out/c15_pytest.log
run() { echo "== $*" | tee -a gpurun_out/c15_kbench.log; env "$@" timeout 300 python tools/kbench.py $CFG $N 4 2>&1 | tail -1 | cut -c1-200 | tee -a gpurun_out/c15_kbench.log; }
CFG=4 N=400000
run CLS_NO_FRAG=1
run CLS_X=1
