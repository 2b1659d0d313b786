#!/bin/bash
# Builds compile-time variants of kernels.cu next to the default library (no GPU needed; ~25 s each):
#   bash tools/build_variants.sh name1:"-DFLAG=..." name2:"..."      ->  classeq2_b200/libclasseq_b200_<name>.so
# then, on the GPU box:  CLASSEQ_B200_LIB=$PWD/classeq2_b200/libclasseq_b200_<name>.so python tools/kbench.py 2 1000000 10
set -e
cd "$(dirname "$0")/../classeq2_b200/csrc"
make -j8
for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    make variant NAME=$name EXTRA="$flags"
done
ls -la ../libclasseq_b200*.so
