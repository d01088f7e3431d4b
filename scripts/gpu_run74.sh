#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "metrics_kernels or training_state or optimize_parameters" 2>&1 | tail -25 | tee gpurun_out/r74_tests.log
timeout 300 python -m selfc_b200.train_loop -opt selfc_b200/configs/selfc_large_train_synthetic.yml --niter 3 2>&1 | tail -5 | tee gpurun_out/r74_train_loop.log
