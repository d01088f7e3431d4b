#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
B1="python bench.py --mode bf16 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-train"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file $O/r2c_launches.csv $B1 > $O/r2c_launches_bench.log 2>&1; echo "launch list rc=$?"
python scripts/launch_summary.py $O/r2c_launches.csv
