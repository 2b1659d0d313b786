#!/bin/bash
# Round 2, call 24 (8 GPUs): the driver's N = 8 command with the default (mixed) packing, with device packing and with host packing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_routed.py tests/test_multi_device.py tests/test_device_pack.py -m gpu -x -q > gpurun_out/c24_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c24_pytest.log
tail -3 gpurun_out/c24_pytest.log
P=29531
for MODE in auto device; do
  P=$((P+1))
  if [ $MODE = auto ]; then EXTRA=""; ENVV="CLS_X=1"; else EXTRA="--no-config5"; ENVV="CLS_PACK=$MODE"; fi
  env $ENVV timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 20 --warmup 5 $EXTRA > gpurun_out/c24_bench_n8_$MODE.json 2> gpurun_out/c24_bench_n8_$MODE.err
  python - $MODE <<'PY'
import json, sys
for l in open(f'gpurun_out/c24_bench_n8_{sys.argv[1]}.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(sys.argv[1], 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']/1e6,1), round(d['e2e']['ms_per_step'],2), d['e2e'].get('pack','')[:12], d['e2e']['breakdown_ms_rank0'], d['e2e'].get('ms_per_step_by_rank'), 'clk', d['clocks'].get('samples'))
        c5=d.get('config5') or {}
        if c5: print('  config5', c5.get('value'), c5.get('ms_per_step'), c5.get('stage_ms_per_step_rank0'), c5.get('error'))
PY
done
