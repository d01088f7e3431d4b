#!/bin/bash
# A/B of an environment knob on 28-frame 1080p runs:  gpurun -- 'bash scripts/gpu_ab.sh SELFC_GA_FORK 1 0'
K=$1; shift
mkdir -p gpurun_out
for v in "$@" "$@"; do
  env $K=$v timeout 300 python bench.py --frames 28 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-train 2>/dev/null | grep '^{' | tail -1 > gpurun_out/ab_${K}_$v.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${K}_$v.json"))
print("$K=$v", round(d["value"],1), "fps", d["clocks"]["sm_mhz"], "MHz", round(d["ms_per_step"],2), "ms/step")
PY
done
