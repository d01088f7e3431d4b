#!/bin/bash
# round-2 profiles: ncu launch list of one GOP + full captures of the two tcgen05 kernels (pair + P2 defaults)
mkdir -p gpurun_out
B="python bench.py --mode bf16 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 440 -c 560 --csv --log-file gpurun_out/r86_launches.csv $B > gpurun_out/r86_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc3 -s 16 -c 8 -o gpurun_out/r86_conv3 $B > gpurun_out/r86_n1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:temporal_tc_kernel -s 16 -c 6 -o gpurun_out/r86_temporal $B > gpurun_out/r86_n2.log 2>&1
ls -la gpurun_out | grep r86
