#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --workload train --steps 1 --warmup 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_kernel -s 40 -c 6 -o gpurun_out/r77_wgrad $B > gpurun_out/r77_n1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2200 -c 2400 --csv --log-file gpurun_out/r77_launches.csv $B > gpurun_out/r77_n2.log 2>&1
ls -la gpurun_out | grep r77
