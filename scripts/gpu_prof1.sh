#!/bin/bash
mkdir -p gpurun_out
# launch list (cold-cache, serialised): one 7-frame GOP, bf16 mode
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1_launches_bf16.csv \
   python bench.py --mode bf16 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r1_launches_bench.log 2>&1
tail -2 gpurun_out/r1_launches_bench.log | cut -c1-300
# full capture of the two tensor-core kernels: launches in block F conv4 (cin 144) and a G conv5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc -s 20 -c 4 -o gpurun_out/r1_prof_conv3x3 \
   python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r1_prof_conv3x3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:temporal_tc -s 4 -c 3 -o gpurun_out/r1_prof_temporal \
   python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r1_prof_temporal.log 2>&1
ls -la gpurun_out/
