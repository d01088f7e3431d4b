#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I selfc_b200/csrc scripts/ubench/mma2_rate.cu -o /tmp/mma2_rate -lcuda 2>&1 | grep -v warning | head -5
timeout 120 /tmp/mma2_rate | tee gpurun_out/ubench_mma2_rate.txt
