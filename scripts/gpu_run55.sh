#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:conv3x3 -c 16 --csv --log-file gpurun_out/r55_conv.csv python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > /dev/null 2>&1
grep conv3x3 gpurun_out/r55_conv.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '
