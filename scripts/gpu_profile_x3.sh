#!/bin/bash
# ncu evidence for the round-2 (second half) kernels:  gpurun --timeout 1500 -- 'bash scripts/gpu_profile_x3.sh r2d'
#   <tag>_x3_conv.ncu-rep      --set full rows of conv3x3_tc3_kernel<PAIR, !P2, X2> (BF16X3 mode, one 1080p GOP: wide layers of an F block)
#   <tag>_x3_temporal.ncu-rep  ... of temporal_tc_kernel<3, false, X2>
#   <tag>_wgrad.ncu-rep        ... of wgrad_tc_kernel and of the input-gradient form of the conv kernel (one training step, 7x256x448)
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
B1="python bench.py --mode bf16x3 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-train --no-gate-mode"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc3" -s 300 -c 6 -o $O/${TAG}_x3_conv $B1 > $O/${TAG}_ncu_x3_conv.log 2>&1; echo "conv rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"temporal_tc" -s 100 -c 4 -o $O/${TAG}_x3_temporal $B1 > $O/${TAG}_ncu_x3_temporal.log 2>&1; echo "temporal rc=$?"
T1="python bench.py --workload train --steps 1 --warmup 1"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"wgrad_tc_kernel" -s 300 -c 5 -o $O/${TAG}_wgrad $T1 > $O/${TAG}_ncu_wgrad.log 2>&1; echo "wgrad rc=$?"
for n in x3_conv x3_temporal wgrad; do python scripts/ncu_summary.py $O/${TAG}_$n.ncu-rep $O/${TAG}_$n.csv; done
ls -la $O/${TAG}_*.csv
