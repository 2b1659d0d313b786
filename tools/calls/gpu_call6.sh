#!/bin/bash
# Round 2, call 6: giant reads + multi-device tests, scan2 after the bookkeeping diet, murmur FMA variant, L2 granularity traffic, bench.py config 3
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_giant_reads.py tests/test_scan2.py tests/test_gpu_parity.py tests/test_routed.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/c6_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c6_pytest.log
tail -8 gpurun_out/c6_pytest.log
for CFG in 2 3; do
  timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c6_kbench.log
  for V in murfma m4; do
    CLASSEQ_B200_LIB=$PWD/classeq2_b200/libclasseq_b200_$V.so timeout 200 python tools/kbench.py $CFG 1000000 10 2>&1 | tail -1 | cut -c1-120 | tee -a gpurun_out/c6_kbench.log
  done
done
CLASSEQ_B200_LIB=$PWD/classeq2_b200/libclasseq_b200_murfma.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hashes or golden or synthetic" 2>&1 | tail -2 | tee gpurun_out/c6_pytest_murfma.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan2_kernel|descend_kernel" -c 2 -o gpurun_out/c6_prof_cfg2 python tools/kbench.py 2 1000000 1 > gpurun_out/c6_ncu_cfg2.log 2>&1
CLS_L2_FETCH=32 timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:"scan2_kernel" -c 1 --csv --log-file gpurun_out/c6_l2fetch32.csv python tools/kbench.py 3 1000000 1 > /dev/null 2>&1
CLS_L2_FETCH=128 timeout 600 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:"scan2_kernel" -c 1 --csv --log-file gpurun_out/c6_l2fetch128.csv python tools/kbench.py 3 1000000 1 > /dev/null 2>&1
tail -2 gpurun_out/c6_l2fetch32.csv gpurun_out/c6_l2fetch128.csv
timeout 900 python bench.py --steps 3 --warmup 3 --cpu-seconds 5 > gpurun_out/c6_bench_cfg3.json 2> gpurun_out/c6_bench_cfg3.err; tail -c 1500 gpurun_out/c6_bench_cfg3.json; tail -3 gpurun_out/c6_bench_cfg3.err
