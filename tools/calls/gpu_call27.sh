#!/bin/bash
# Round 2, call 27: the driver's one-GPU bench command (config 3 + nested config 4), wall time of the run
mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/c27_bench_cfg3.json 2> gpurun_out/c27_bench_cfg3.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/c27_bench_cfg3.err
python - <<'PY'
import json
for l in open('gpurun_out/c27_bench_cfg3.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e6,1), round(d['e2e']['ms_per_step'],2), d['e2e']['pack'][:6], 'frac', round(d['roofline']['frac'],4), 'traffic', d['roofline']['traffic'], 'parity', d['parity'], 'clocks', d['clocks'])
        print('secondary', json.dumps(d['roofline']['secondary'])[:900])
        print('config4', json.dumps(d.get('config4'))[:1200])
PY
