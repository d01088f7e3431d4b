#!/bin/bash
# 2-GPU check of the default bench under torchrun (rescale + train)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r87_bench_2gpu.json; cut -c1-330 gpurun_out/r87_bench_2gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --steps 5 --warmup 3 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r87_train_2gpu.json; cut -c1-260 gpurun_out/r87_train_2gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl reference --steps 1 --warmup 1 2>/dev/null | grep '^{' | tail -1 | cut -c1-200
