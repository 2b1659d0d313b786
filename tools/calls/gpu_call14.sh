#!/bin/bash
# Round 2, call 14 (2 GPUs): multi-GPU tests, bench under torchrun with device packing (what an 8-GPU box picks), nested config 5
mkdir -p gpurun_out
nvidia-smi -L | head -3; nproc
timeout 900 python -m pytest tests/test_routed.py tests/test_multi_device.py -m gpu -x -q > gpurun_out/c14_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/c14_pytest.log
tail -5 gpurun_out/c14_pytest.log
CLS_PACK=device timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/c14_bench_n2_device.json 2> gpurun_out/c14_bench_n2_device.err; tail -c 3000 gpurun_out/c14_bench_n2_device.json; tail -3 gpurun_out/c14_bench_n2_device.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-config5 > gpurun_out/c14_bench_n2_auto.json 2> gpurun_out/c14_bench_n2_auto.err; tail -c 1200 gpurun_out/c14_bench_n2_auto.json | head -c 800; tail -3 gpurun_out/c14_bench_n2_auto.err
