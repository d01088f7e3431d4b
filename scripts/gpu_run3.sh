#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16 or drop_in" 2>&1 | tail -40 > gpurun_out/r5_bf16.log
tail -25 gpurun_out/r5_bf16.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 600 python bench.py --mode bf16 --frames 14 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r5_bench_bf16.json 2> gpurun_out/r5_bench_bf16.err; tail -c 1500 gpurun_out/r5_bench_bf16.json; tail -3 gpurun_out/r5_bench_bf16.err
