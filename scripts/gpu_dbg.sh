#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused_dense_block or variants" 2>&1 | tail -2
timeout 300 python scripts/dbg_fused.py 2>&1 | grep "per step" | cut -c1-330
timeout 300 python bench.py --frames 28 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-train 2>/dev/null | grep '^{' | tail -1 > gpurun_out/r2c_b28.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2c_b28.json"))
print(round(d["value"],1), "fps", d["clocks"], {k:v["ms"] for k,v in d["roofline"]["classes"].items()}, "frac", round(d["roofline"]["frac"],3))
PY
