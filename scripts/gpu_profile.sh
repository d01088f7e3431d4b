#!/bin/bash
# ncu evidence for one 1080p bf16 GOP:  gpurun --timeout 1800 -- 'bash scripts/gpu_profile.sh r2c'
#   <tag>_launches.csv          every launch: device time + DRAM bytes (cold caches, serialised; compare SHARES)  -> scripts/launch_summary.py
#   <tag>_dense_fused.ncu-rep   --set full capture of the fused dense-block kernel                                -> scripts/ncu_summary.py
#   <tag>_dbg_fused.txt         per-role barrier-wait cycles of the fused kernel (SELFC_TC_DBG=1 build of the kernel is selected at run time)
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
B1="python bench.py --mode bf16 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-train"
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 2500 --csv --log-file $O/${TAG}_launches.csv $B1 > $O/${TAG}_launches_bench.log 2>&1; echo "launch list rc=$?"
python scripts/launch_summary.py $O/${TAG}_launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"dense_fused" -s 40 -c 6 -o $O/${TAG}_dense_fused $B1 > $O/${TAG}_ncu_full.log 2>&1; echo "full capture rc=$?"
timeout 300 python scripts/dbg_fused.py > $O/${TAG}_dbg_fused.txt 2>&1; grep -A1 "schedule" $O/${TAG}_dbg_fused.txt | cut -c1-330
