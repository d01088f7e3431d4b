#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "variants" 2>&1 | tail -12 | tee gpurun_out/r81_tests.log
