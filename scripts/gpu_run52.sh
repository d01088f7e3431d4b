#!/bin/bash
for i in $(seq 1 14); do timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -s -k "fa_ or quantize or dense_block_vs or global_agg_vs or gmm_sample" 2>&1 | grep -E "mismatch|passed|failed" | tail -2; done
