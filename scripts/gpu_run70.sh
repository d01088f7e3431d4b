#!/bin/bash
# CTA-pair conv3x3: parity tests, then A/B bench against the one-CTA kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv3x3 or bf16 or dense_block" 2>&1 | tail -15 > gpurun_out/r70_tests.log
cat gpurun_out/r70_tests.log
for pair in 1 0; do
  SELFC_TC3_PAIR=$pair timeout 600 python bench.py --steps 3 --warmup 3 --frames 28 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/r70_bench_pair$pair.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r70_bench_pair$pair.json"))
print("pair=$pair", round(d["value"],1), "fps", {k:v["ms"] for k,v in d["roofline"]["classes"].items()}, d["clocks"])
PY
done
