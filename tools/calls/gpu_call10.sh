#!/bin/bash
# Round 2, call 10: whole GPU suite on the committed state, random-probe micro-benchmark (rate + DRAM bytes per probe by table size)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c10_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c10_pytest.log
tail -4 gpurun_out/c10_pytest.log
timeout 300 tools/micro/probe_bench > gpurun_out/c10_probe_bench.log 2>&1; cat gpurun_out/c10_probe_bench.log
timeout 600 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --csv --log-file gpurun_out/c10_probe_ncu.csv tools/micro/probe_bench 64 128 256 1024 > /dev/null 2>&1
grep -c probe_kernel gpurun_out/c10_probe_ncu.csv
