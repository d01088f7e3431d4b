#!/bin/bash
# 2-GPU line of the session-4 build (rescale, default bench under torchrun)
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/r97_err.log | grep '^{' | tail -1 > gpurun_out/r97_bench_2gpu.json; cut -c1-200 gpurun_out/r97_bench_2gpu.json; tail -2 gpurun_out/r97_err.log
