#!/bin/bash
# Round 2, call 23: chunk ramp A/B (growth factor, first chunk) on the end-to-end time of 10 M reads
mkdir -p gpurun_out
run() { env "$@" PIN=1 timeout 300 python tools/e2e_bench.py 10000000 3 2>&1 | tail -1 | cut -c1-330 | sed 's/status_hist.*sum(node)=[0-9]* //' | tee -a gpurun_out/c23_e2e.log; }
run CLS_X=1
run CLS_CHUNK_GROWTH=150
run CLS_CHUNK_GROWTH=130
run CLS_CHUNK_GROWTH=150 CLS_CHUNK_FIRST=8
run CLS_CHUNK_GROWTH=150 CLS_CHUNK_FIRST=8 CLS_CHUNK_RAMP=3
run CLS_CHUNK_GROWTH=150 CLS_CHUNK_MBASES=90
run CLS_CHUNK_GROWTH=150 CLS_HOST_THREADS=12
