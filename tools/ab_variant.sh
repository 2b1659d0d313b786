#!/bin/bash
# A/B of a compile-time kernel variant against the default library, in ONE gpurun call:
#   HERE (no GPU):   make -C classeq2_b200/csrc variant NAME=plain EXTRA=-DCLS_INSERT_PLAIN=1
#   on the GPU box:  bash tools/ab_variant.sh plain [config=2] [n_reads=1000000]
# 1. the placement parity tests with the variant library (bit-exactness comes first), 2. kbench of both libraries
# (status histogram and checksums of the results are printed: they must be identical), logs in gpurun_out/.
NAME=${1:?variant name}
CFG=${2:-2}
N=${3:-1000000}
LIB=$PWD/classeq2_b200/libclasseq_b200_$NAME.so
mkdir -p gpurun_out
[ -f "$LIB" ] || { echo "missing $LIB: build it first (make -C classeq2_b200/csrc variant NAME=$NAME EXTRA=...)"; exit 2; }
CLASSEQ_B200_LIB=$LIB timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_routed.py -m gpu -x -q > gpurun_out/ab_${NAME}_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/ab_${NAME}_pytest.log
tail -3 gpurun_out/ab_${NAME}_pytest.log
{ timeout 120 python tools/kbench.py $CFG $N 10; CLASSEQ_B200_LIB=$LIB timeout 120 python tools/kbench.py $CFG $N 10; } > gpurun_out/ab_${NAME}_kbench.log 2>&1
cat gpurun_out/ab_${NAME}_kbench.log
