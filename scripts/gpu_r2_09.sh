#!/bin/bash
mkdir -p gpurun_out
for v in spin hint; do
  cp selfc_b200/libselfc_b200.so.$v selfc_b200/libselfc_b200.so
  for fr in 28 100; do
  timeout 300 python bench.py --frames $fr --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2b_09_err.log | grep '^{' | tail -1 > gpurun_out/r2b_09_$v$fr.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r2b_09_$v$fr.json"))
print("$v frames=$fr", round(d["value"],1), "fps", d["clocks"], {k:v["ms"] for k,v in d["roofline"]["classes"].items()}, "frac", round(d["roofline"]["frac"],3))
PY
  done
done
