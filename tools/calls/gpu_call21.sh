#!/bin/bash
# Round 2, call 21: the whole GPU suite, smoke, bench lines (configs 3, 2, 4), launch list and full captures of the final kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c21_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c21_pytest.log
tail -4 gpurun_out/c21_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c21_bench_cfg3.json 2> gpurun_out/c21_bench_cfg3.err; tail -c 300 gpurun_out/c21_bench_cfg3.json; tail -2 gpurun_out/c21_bench_cfg3.err
timeout 600 python bench.py --config 2 --steps 10 --warmup 3 --cpu-seconds 5 > gpurun_out/c21_bench_cfg2.json 2> gpurun_out/c21_bench_cfg2.err; tail -c 200 gpurun_out/c21_bench_cfg2.json
timeout 900 python bench.py --config 4 --steps 3 --warmup 3 --cpu-seconds 5 > gpurun_out/c21_bench_cfg4.json 2> gpurun_out/c21_bench_cfg4.err; tail -c 200 gpurun_out/c21_bench_cfg4.json; tail -2 gpurun_out/c21_bench_cfg4.err
for V in 125 250 500; do CLS_CHUNK_MBASES=$V PIN=1 timeout 300 python tools/e2e_bench.py 10000000 3 2>&1 | tail -1 | cut -c1-200 | tee -a gpurun_out/c21_e2e_chunks.log; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c21_launches_bench.csv python bench.py --steps 2 --warmup 1 --cpu-seconds 1 --reads 2000000 > gpurun_out/c21_ncu_bench.log 2>&1
for CFG in 2 3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scan2_kernel|descend" -c 2 -o gpurun_out/c21_prof_cfg$CFG python tools/kbench.py $CFG 1000000 1 > gpurun_out/c21_ncu_cfg$CFG.log 2>&1
done
ls -la gpurun_out/c21_prof_cfg*.ncu-rep
