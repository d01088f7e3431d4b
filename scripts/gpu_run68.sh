#!/bin/bash
# f2: uint8 ingest/egress tests + bench with the u8 e2e leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "u8 or drop_in" 2>&1 | tail -15 > gpurun_out/r68_tests.log
cat gpurun_out/r68_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --frames 28 2>&1 | tail -1 > gpurun_out/r68_bench.json
cat gpurun_out/r68_bench.json
