#!/bin/bash
# DRAM bytes per kernel class of one GOP with and without the zigzag tile order
mkdir -p gpurun_out
B="python bench.py --mode bf16 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
for zz in 1 0; do
SELFC_ZIGZAG=$zz timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none --launch-skip 440 -c 560 --csv --log-file gpurun_out/r72_launches_z$zz.csv $B > gpurun_out/r72_z$zz.log 2>&1
done
ls -la gpurun_out | grep r72
