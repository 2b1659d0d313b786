#!/bin/bash
# One GPU call: the device model builder's tests first (their own log), then the whole GPU suite, then a timing
# of the two builders.  Every step has its own timeout; logs go to gpurun_out/.
mkdir -p gpurun_out
timeout 45 python -m pytest tests/test_zbuild_device.py -m gpu -q > gpurun_out/zbuild.log 2>&1; echo "rc=$?" >> gpurun_out/zbuild.log
tail -3 gpurun_out/zbuild.log
timeout 90 python -m pytest tests -m gpu -x -q --deselect tests/test_zbuild_device.py --durations=8 > gpurun_out/pytest_gpu_final.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -12 gpurun_out/pytest_gpu_final.log
timeout 20 python tools/build_bench.py 1000 1000 > gpurun_out/build_bench.log 2>&1; echo "rc=$?" >> gpurun_out/build_bench.log
cat gpurun_out/build_bench.log
