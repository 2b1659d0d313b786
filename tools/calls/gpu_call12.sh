#!/bin/bash
# Round 2, call 12: device-side 2-bit packing (tests + end-to-end A/B against the host packer, with all host threads and with 4)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_device_pack.py tests/test_abi.py -m gpu -x -q > gpurun_out/c12_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c12_pytest.log
tail -15 gpurun_out/c12_pytest.log
run() { env "$@" timeout 300 python tools/e2e_bench.py $N 3 2>&1 | tail -1 | cut -c1-330 | tee -a gpurun_out/c12_e2e.log; }
N=1250000
run CLS_PACK=host
run CLS_PACK=device
run CLS_PACK=device PIN=1
run CLS_PACK=host CLS_HOST_THREADS=4
run CLS_PACK=device CLS_HOST_THREADS=4
run CLS_PACK=device PIN=1 CLS_HOST_THREADS=4
N=10000000
run CLS_PACK=host
run CLS_PACK=device PIN=1
run CLS_PACK=device PIN=1 CLS_HOST_THREADS=4
