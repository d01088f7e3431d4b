#!/bin/bash
# ncu --set full captures: conv3x3_tc3 (F block conv1..4), temporal kernel variants (conv5 F/H/G, STP conv5, GA apply, head), sampler
mkdir -p gpurun_out
B="python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc3 -s 20 -c 4 -o gpurun_out/r22_conv3 $B > gpurun_out/r22_conv3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:temporal_tc -s 24 -c 3 -o gpurun_out/r22_temporal_inv $B > gpurun_out/r22_t1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:temporal_tc -s 34 -c 7 -o gpurun_out/r22_temporal_stp $B > gpurun_out/r22_t2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gmm_sample -c 1 -o gpurun_out/r22_sampler $B > gpurun_out/r22_s.log 2>&1
ls -la gpurun_out | grep r22
