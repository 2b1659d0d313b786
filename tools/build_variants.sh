#!/bin/bash
# Builds the prepared compile-time kernel variants (no GPU needed; about 20 s each) next to the default library:
#   plain     warp-owned tables without shared-memory atomics          (-DCLS_INSERT_PLAIN)
#   d8 / d4   descend_kernel asked for 8 / 4 resident CTAs per SM      (-DCLS_DESCEND_MINB)
#   plain_d8  both
#   scan5     scan_kernel asked for 5 resident CTAs per SM             (-DCLS_SCAN_MINB=5)
# then, on the GPU box:  bash tools/ab_variant.sh plain d8 d4 plain_d8 scan5
set -e
cd "$(dirname "$0")/../classeq2_b200/csrc"
make -j8
make variant NAME=plain EXTRA=-DCLS_INSERT_PLAIN=1
make variant NAME=d8 EXTRA=-DCLS_DESCEND_MINB=8
make variant NAME=d4 EXTRA=-DCLS_DESCEND_MINB=4
make variant NAME=plain_d8 EXTRA="-DCLS_INSERT_PLAIN=1 -DCLS_DESCEND_MINB=8"
make variant NAME=scan5 EXTRA=-DCLS_SCAN_MINB=5
ls -la ../libclasseq_b200*.so
