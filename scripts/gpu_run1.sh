#!/bin/bash
# first GPU call: parity tests, smoke, short fp32 bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r1_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r1_pytest.log
cat gpurun_out/r1_pytest.log | tail -15
timeout 300 python __graft_entry__.py smoke > gpurun_out/r1_smoke.log 2>&1; tail -5 gpurun_out/r1_smoke.log
timeout 600 python bench.py --mode fp32 --frames 14 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1_bench_fp32.json 2> gpurun_out/r1_bench_fp32.err; tail -c 3000 gpurun_out/r1_bench_fp32.json; tail -5 gpurun_out/r1_bench_fp32.err
