#!/bin/bash
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gmm_sample" 2>&1 | tail -2; done
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6
