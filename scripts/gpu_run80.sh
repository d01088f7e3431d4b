#!/bin/bash
# hybrid TMA + cp.async activation loads in conv3x3: parity, then A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv3x3 or bf16 or dense_block or 1080p" 2>&1 | tail -15 > gpurun_out/r80_tests.log
cat gpurun_out/r80_tests.log
for hy in 1 0; do
  SELFC_TC3_HYBRID=$hy timeout 600 python bench.py --steps 3 --warmup 3 --frames 28 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/r80_bench_h$hy.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r80_bench_h$hy.json"))
print("hybrid=$hy", round(d["value"],1), "fps", {k:v["ms"] for k,v in d["roofline"]["classes"].items()}, d["clocks"]["sm_mhz"])
PY
done
