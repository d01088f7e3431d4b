#!/bin/bash
# full-size agreement tests (1080p, 4K), 2-GPU bench through torchrun, Vid4-shape and 4K-shape throughput
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "full_size" 2>&1 | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r50_bench_2gpu.json 2> gpurun_out/r50_2gpu.err; tail -2 gpurun_out/r50_2gpu.err; cut -c1-260 gpurun_out/r50_bench_2gpu.json
timeout 600 python bench.py --mode bf16 --height 576 --width 704 --frames 105 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r50_bench_vid4shape_bf16.json 2>/dev/null; cut -c1-200 gpurun_out/r50_bench_vid4shape_bf16.json
timeout 600 python bench.py --mode fp32 --height 576 --width 704 --frames 14 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r50_bench_vid4shape_fp32.json 2>/dev/null; cut -c1-200 gpurun_out/r50_bench_vid4shape_fp32.json
timeout 600 python bench.py --mode bf16 --height 2160 --width 3840 --frames 28 --steps 2 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r50_bench_4k_bf16.json 2>/dev/null; cut -c1-200 gpurun_out/r50_bench_4k_bf16.json
