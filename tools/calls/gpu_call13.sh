#!/bin/bash
# Round 2, call 13: just-in-time plan + chunk ramp-down + device pack: tests, end-to-end A/B, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_pack.py tests/test_gpu_parity.py tests/test_multi_device.py tests/test_zwriter_e2e.py -m gpu -x -q > gpurun_out/c13_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c13_pytest.log
tail -15 gpurun_out/c13_pytest.log
run() { env "$@" timeout 300 python tools/e2e_bench.py $N 3 2>&1 | tail -1 | cut -c1-330 | tee -a gpurun_out/c13_e2e.log; }
N=1250000
run CLS_PACK=host
run CLS_PACK=host CLS_NO_FASTPLAN=1
run CLS_PACK=device PIN=1
run CLS_PACK=host CLS_HOST_THREADS=4
run CLS_PACK=device PIN=1 CLS_HOST_THREADS=4
run CLS_PACK=device PIN=1 CLS_HOST_THREADS=4 CLS_NO_FASTPLAN=1
run CLS_PACK=device CLS_HOST_THREADS=4
N=10000000
run CLS_PACK=host
run CLS_PACK=device PIN=1
timeout 900 python bench.py --steps 5 --warmup 3 --cpu-seconds 5 > gpurun_out/c13_bench_cfg3.json 2> gpurun_out/c13_bench_cfg3.err; tail -c 2500 gpurun_out/c13_bench_cfg3.json; tail -3 gpurun_out/c13_bench_cfg3.err
