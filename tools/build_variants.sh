#!/bin/bash
# Builds the prepared compile-time kernel variants (no GPU needed; about 20 s each) next to the default library:
#   plain     warp-owned tables without shared-memory atomics          (-DCLS_INSERT_PLAIN)
#   d8 / d4   descend_kernel asked for 8 / 4 resident CTAs per SM      (-DCLS_DESCEND_MINB)
#   plain_d8  both
#   scan5     scan_kernel asked for 5 resident CTAs per SM             (-DCLS_SCAN_MINB=5)
#   dyn       reads handed out in blocks of 4 from a global counter    (-DCLS_DYNAMIC_READS)
#   all       plain + d8 + dyn
# then, on the GPU box:  bash tools/ab_variant.sh plain d8 d4 plain_d8 scan5 dyn all
# (the host-side experiment CLS_PIPE=3 needs no build: CLS_PIPE=3 python tools/e2e_bench.py)
set -e
cd "$(dirname "$0")/../classeq2_b200/csrc"
make -j8
make variant NAME=plain EXTRA=-DCLS_INSERT_PLAIN=1
make variant NAME=d8 EXTRA=-DCLS_DESCEND_MINB=8
make variant NAME=d4 EXTRA=-DCLS_DESCEND_MINB=4
make variant NAME=plain_d8 EXTRA="-DCLS_INSERT_PLAIN=1 -DCLS_DESCEND_MINB=8"
make variant NAME=scan5 EXTRA=-DCLS_SCAN_MINB=5
make variant NAME=dyn EXTRA=-DCLS_DYNAMIC_READS=1
make variant NAME=all EXTRA="-DCLS_INSERT_PLAIN=1 -DCLS_DESCEND_MINB=8 -DCLS_DYNAMIC_READS=1"
ls -la ../libclasseq_b200*.so
