#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 20 -c 4 -o gpurun_out/r3_prof_conv3x3 \
   python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r3_prof_conv3x3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:temporal_tc -s 0 -c 3 -o gpurun_out/r3_prof_temporal \
   python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r3_prof_temporal.log 2>&1
