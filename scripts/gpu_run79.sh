#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --workload train --steps 1 --warmup 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_simt_kernel -s 300 -c 12 -o gpurun_out/r79_simt $B > gpurun_out/r79_n1.log 2>&1
ls -la gpurun_out | grep r79
