#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for g in 1 4 10; do
timeout 300 python bench.py --height 576 --width 704 --frames 700 --gops-per-launch $g --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>$O/r2b_15_err.log | grep '^{' | tail -1 > $O/r2b_15_vid4_g$g.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2b_15_vid4_g$g.json"))
print("gpl=$g", round(d["value"],1), "fps", d["clocks"], {k:v["ms"] for k,v in d["roofline"]["classes"].items()}, "frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"])
PY
done
