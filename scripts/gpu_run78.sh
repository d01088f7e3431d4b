#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "backward or train_step or optimizer or optimize_parameters or training_state or vid4 or network_vs_golden" 2>&1 | tail -8 | tee gpurun_out/r78_tests.log
timeout 600 python bench.py --workload train --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r78_train.json; cut -c1-260 gpurun_out/r78_train.json
timeout 600 python bench.py --mode fp32 --height 576 --width 704 --frames 28 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r78_fp32.json; cut -c1-200 gpurun_out/r78_fp32.json
