#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "backward or train_step or optimizer or optimize_parameters or training_state" 2>&1 | tail -8 | tee gpurun_out/r76_tests.log
timeout 600 python bench.py --workload train --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r76_train.json; cut -c1-260 gpurun_out/r76_train.json
