#!/bin/bash
# Round 2, call 7: register budget picked by table size; e2e pipeline A/B (three streams default, chunk sizes, read blocks)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scan2.py tests/test_gpu_parity.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/c7_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c7_pytest.log
tail -3 gpurun_out/c7_pytest.log
for CFG in 2 3; do timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c7_kbench.log; done
{
timeout 200 python tools/e2e_bench.py 1000000 2
CLS_PIPE=2 timeout 200 python tools/e2e_bench.py 1000000 2
CLS_CHUNK_MBASES=12 timeout 200 python tools/e2e_bench.py 1000000 2
CLS_CHUNK_MBASES=48 timeout 200 python tools/e2e_bench.py 1000000 2
timeout 200 python tools/e2e_bench.py 1250000 3
CLS_PIPE=2 timeout 200 python tools/e2e_bench.py 1250000 3
timeout 300 python tools/e2e_bench.py 10000000 3
CLS_PIPE=2 timeout 300 python tools/e2e_bench.py 10000000 3
CLS_CHUNK_MBASES=24 timeout 300 python tools/e2e_bench.py 10000000 3
CLS_CHUNK_MBASES=250 timeout 300 python tools/e2e_bench.py 10000000 3
timeout 300 python tools/e2e_bench.py 10000000 3 2
CLS_DEBUG_TIMING=1 timeout 200 python tools/e2e_bench.py 1250000 3 2>&1 | tail -12
} 2>&1 | cut -c1-400 | tee gpurun_out/c7_e2e.log
