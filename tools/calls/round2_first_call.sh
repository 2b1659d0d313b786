#!/bin/bash
# The first GPU call of the next round, in one go (about 12 minutes of box time on one B200):
#   HERE (no GPU), before the call:   bash tools/build_variants.sh
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# 1. the GPU tests written after round 1's GPU time was spent, then the whole GPU suite;
# 2. parity + kernel timing of every prepared kernel variant against the default library;
# 3. end-to-end cls_place_batch with the default pipeline and with CLS_PIPE=3, default and best-looking variants;
# 4. the device model builder's timing.  Logs in gpurun_out/.
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_zwriter_e2e.py tests/test_place_sequences.py -m gpu -q > gpurun_out/r2_new_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_new_tests.log
tail -3 gpurun_out/r2_new_tests.log
timeout 200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2_pytest_gpu.log
tail -3 gpurun_out/r2_pytest_gpu.log
bash tools/ab_variant.sh plain d8 d4 plain_d8 scan5 dyn all 2>&1 | tee gpurun_out/r2_ab.log | grep -v "^$" | tail -40
{ timeout 100 python tools/e2e_bench.py; CLS_PIPE=3 timeout 100 python tools/e2e_bench.py;
  for v in dyn all; do L=$PWD/classeq2_b200/libclasseq_b200_$v.so; [ -f $L ] && { CLASSEQ_B200_LIB=$L timeout 100 python tools/e2e_bench.py; CLASSEQ_B200_LIB=$L CLS_PIPE=3 timeout 100 python tools/e2e_bench.py; }; done; } 2>&1 | tee gpurun_out/r2_e2e.log
timeout 60 python tools/build_bench.py 1000 1000 2>&1 | tee gpurun_out/r2_build_bench.log
