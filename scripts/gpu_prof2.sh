#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:temporal_tc -c 80 --csv --log-file gpurun_out/r2_launches_temporal.csv \
   python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:temporal_tc -s 2 -c 1 -o gpurun_out/r2_prof_temporal_y2 \
   python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r2_prof_temporal.log 2>&1
ls -la gpurun_out | tail -5
