#!/bin/bash
# A/B of compile-time kernel variants against the default library, in ONE gpurun call:
#   HERE (no GPU):   bash tools/build_variants.sh        (or: make -C classeq2_b200/csrc variant NAME=... EXTRA=...)
#   on the GPU box:  [CFG=2] [N=1000000] bash tools/ab_variant.sh plain d8 ...
# For every variant: 1. the placement parity tests with the variant library (bit-exactness comes first), 2. kbench
# (status histogram and checksums of the results are printed: they must equal the default library's); logs in gpurun_out/.
CFG=${CFG:-2}
N=${N:-1000000}
mkdir -p gpurun_out
timeout 120 python tools/kbench.py $CFG $N 10 2>&1 | tee gpurun_out/ab_default_kbench.log
for NAME in "$@"; do
    LIB=$PWD/classeq2_b200/libclasseq_b200_$NAME.so
    [ -f "$LIB" ] || { echo "missing $LIB: build it first (bash tools/build_variants.sh)"; continue; }
    CLASSEQ_B200_LIB=$LIB timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_routed.py -m gpu -x -q > gpurun_out/ab_${NAME}_pytest.log 2>&1
    echo "rc=$?" >> gpurun_out/ab_${NAME}_pytest.log
    echo "== $NAME: $(tail -2 gpurun_out/ab_${NAME}_pytest.log | tr '\n' ' ')"
    CLASSEQ_B200_LIB=$LIB timeout 120 python tools/kbench.py $CFG $N 10 2>&1 | tee gpurun_out/ab_${NAME}_kbench.log
done
