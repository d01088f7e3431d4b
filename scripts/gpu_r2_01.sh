#!/bin/bash
# round 2, run 1: A-operand-in-TMEM semantics + rate
mkdir -p gpurun_out
timeout 120 ./build/ubench/mma_tmem_a > gpurun_out/r2b_ubench_mma_tmem_a.txt 2>&1; echo rc=$?; cat gpurun_out/r2b_ubench_mma_tmem_a.txt
