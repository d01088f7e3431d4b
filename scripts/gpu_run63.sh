#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload train --steps 3 --warmup 1 > gpurun_out/r63_train_1gpu.json 2> gpurun_out/r63_err.txt; tail -3 gpurun_out/r63_err.txt; cut -c1-400 gpurun_out/r63_train_1gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --workload train --gpus 2 --steps 3 --warmup 1 2> gpurun_out/r63_err2.txt | grep '^{' > gpurun_out/r63_train_2gpu.json; tail -3 gpurun_out/r63_err2.txt; cut -c1-400 gpurun_out/r63_train_2gpu.json
