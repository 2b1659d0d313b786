#!/bin/bash
# Builder timing on config 3's model (10 000 tips x 600 bp), then the ncu launch list of one device build of config 2's model.
mkdir -p gpurun_out
timeout 28 python tools/build_bench.py 10000 600 1002 > gpurun_out/build_bench_cfg3.log 2>&1; echo "rc=$?" >> gpurun_out/build_bench_cfg3.log
cat gpurun_out/build_bench_cfg3.log
timeout 25 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_build.csv python tools/build_bench.py 1000 1000 1001 device-only > gpurun_out/ncu_build.log 2>&1; echo "rc=$?" >> gpurun_out/ncu_build.log
tail -3 gpurun_out/ncu_build.log
