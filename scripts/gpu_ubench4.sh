#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I selfc_b200/csrc scripts/ubench/tma_rate.cu -o /tmp/tma_rate -lcuda 2>&1 | grep -v warning | head -5
timeout 120 /tmp/tma_rate | tee gpurun_out/ubench_tma_rate.txt
