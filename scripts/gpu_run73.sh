#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "2x_operators or fa_ or u8" 2>&1 | tail -8 | tee gpurun_out/r73_tests.log
