#!/bin/bash
# Round 2, call 28 (8 GPUs): hash-sharded index with the probe kernels walking the senders in staggered order; 128-byte aligned reply segments
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_routed.py -m gpu -x -q -k "nccl or all_gpus or sharded" > gpurun_out/c28_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c28_pytest.log
tail -3 gpurun_out/c28_pytest.log
P=29551
run() {
  NAME=$1; shift
  P=$((P+1))
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --config 5 --reads 10000000 --steps 3 --warmup 2 --device-build > gpurun_out/c28_bench_cfg5_$NAME.json 2> gpurun_out/c28_bench_cfg5_$NAME.err
  python - $NAME <<'PY'
import json, sys
ok=False
for l in open(f'gpurun_out/c28_bench_cfg5_{sys.argv[1]}.json'):
    if l.startswith('{'):
        d=json.loads(l); ok=True
        print(sys.argv[1], 'value', round(d['value']/1e6,1), 'ms', round(d['ms_per_step'],2), d.get('stage_ms_per_step_rank0'), 'nvlink out GB/s', round((d.get('nvlink') or {}).get('achieved_gbs_out_rank0',0),1), 'parity', d.get('parity'))
if not ok: print(sys.argv[1], 'NO LINE'); print(open(f'gpurun_out/c28_bench_cfg5_{sys.argv[1]}.err').read()[-1500:])
PY
}
run stagger_align32 CLS_X=1
run stagger_align4 CLS_SEG_ALIGN=4
