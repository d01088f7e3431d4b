#!/bin/bash
mkdir -p gpurun_out
for f in 1; do
  SELFC_DB_FUSED=$f timeout 300 python bench.py --frames 28 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2b_05_err$f.log | grep '^{' | tail -1 > gpurun_out/r2b_05_fused$f.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r2b_05_fused$f.json"))
print("fused=$f", round(d["value"],1), "fps", d["clocks"], {k:(v["ms"],v["launches"]) for k,v in d["roofline"]["classes"].items()}, "frac", round(d["roofline"]["frac"],3))
PY
done
