#!/bin/bash
# A/B of an environment knob on the training step (1 and 4 septuplets per step):  gpurun -- 'bash scripts/gpu_ab_train.sh SELFC_TRAIN_PDL 1 0'
K=$1; shift
mkdir -p gpurun_out
for b in 1 4; do
  for v in "$@" "$@"; do
    env $K=$v timeout 300 python bench.py --workload train --septuplets $b --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 > gpurun_out/abt_${K}_${v}_b$b.json
    python - <<PY
import json
d=json.load(open("gpurun_out/abt_${K}_${v}_b$b.json"))
print("$K=$v b=$b", round(d["ms_per_step"],2), "ms/step", round(d["value"],2), d["unit"], d.get("clocks",{}).get("sm_mhz"), "MHz", d.get("gpu_launches"), "launches")
PY
  done
done
