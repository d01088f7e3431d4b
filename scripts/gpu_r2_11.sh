#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused_dense_block" 2>&1 | tail -3
timeout 300 python scripts/dbg_fused.py > gpurun_out/r2b_11_dbg_fused.txt 2>&1; echo rc=$?; grep "per step" gpurun_out/r2b_11_dbg_fused.txt
for fr in 28 100; do
timeout 300 python bench.py --frames $fr --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2b_11_err.log | grep '^{' | tail -1 > gpurun_out/r2b_11_b$fr.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2b_11_b$fr.json"))
print("frames=$fr", round(d["value"],1), "fps", d["clocks"], {k:v["ms"] for k,v in d["roofline"]["classes"].items()}, "frac", round(d["roofline"]["frac"],3))
PY
done
