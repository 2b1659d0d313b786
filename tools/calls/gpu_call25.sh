#!/bin/bash
# Round 2, call 25 (8 GPUs): host threads = this rank's share of the cores; device against mixed packing, and the old 2x share
mkdir -p gpurun_out
P=29541
run() {
  NAME=$1; shift
  P=$((P+1))
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 20 --warmup 5 --no-config5 > gpurun_out/c25_bench_n8_$NAME.json 2> gpurun_out/c25_bench_n8_$NAME.err
  python - $NAME <<'PY'
import json, sys
for l in open(f'gpurun_out/c25_bench_n8_{sys.argv[1]}.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(sys.argv[1], 'value', round(d['value']/1e6,1), 'e2e', round(d['e2e']['value']/1e6,1), round(d['e2e']['ms_per_step'],2), d['e2e'].get('pack','')[:8], d['e2e']['breakdown_ms_rank0'], d['e2e'].get('ms_per_step_by_rank'))
PY
}
run device CLS_PACK=device
run mixed CLS_PACK=mixed
run device_t8 CLS_PACK=device CLS_HOST_THREADS=8
run host CLS_PACK=host
