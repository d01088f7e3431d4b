#!/bin/bash
mkdir -p gpurun_out
for g in 1 2 3; do
  timeout 600 python bench.py --steps 3 --warmup 3 --gops-per-launch $g --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r88_gpl$g.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r88_gpl$g.json"))
print("gops_per_launch=$g", round(d["value"],1), "fps", d["clocks"]["sm_mhz"], d["gpu_launches"])
PY
done
