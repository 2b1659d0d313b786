#!/bin/bash
# Round 2, call 19 (8 GPUs): multi-GPU tests, the driver's N = 8 bench command (config 3 + nested config 5), in-process multi-GPU e2e
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
timeout 600 python -m pytest tests/test_routed.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/c19_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c19_pytest.log
tail -4 gpurun_out/c19_pytest.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c19_bench_n8.json 2> gpurun_out/c19_bench_n8.err; tail -c 600 gpurun_out/c19_bench_n8.err
python - <<'PY'
import json
for l in open('gpurun_out/c19_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('value', d['value']/1e6, 'ms', d['ms_per_step'], 'e2e', d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e'].get('pack'), d['e2e']['breakdown_ms_rank0'])
        c5=d.get('config5') or {}
        print('config5', {k: c5.get(k) for k in ('value','ms_per_step','stage_ms_per_step_rank0','error')}, (c5.get('nvlink') or {}).get('achieved_gbs_out_rank0'))
PY
CLS_PACK=host timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --no-config5 > gpurun_out/c19_bench_n8_hostpack.json 2> gpurun_out/c19_bench_n8_hostpack.err
python - <<'PY'
import json
for l in open('gpurun_out/c19_bench_n8_hostpack.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('hostpack: value', d['value']/1e6, 'e2e', d['e2e']['value']/1e6, d['e2e']['ms_per_step'], d['e2e'].get('pack'), d['e2e']['breakdown_ms_rank0'])
PY
