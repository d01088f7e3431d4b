#!/bin/bash
# round 2: ncu launch list of the 1080p bf16 GOP (fused dense blocks) + full capture of the fused kernel and the temporal kernel
mkdir -p gpurun_out
O=gpurun_out
B1="python bench.py --mode bf16 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file $O/r2b_launches.csv $B1 > $O/r2b_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dense_fused" -s 43 -c 6 -o $O/r2b_dense_fused $B1 > $O/r2b_n1.log 2>&1; echo "full capture rc=$?"
ls -la $O | grep r2b_ | tail -5
