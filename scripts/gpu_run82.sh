#!/bin/bash
# position-pair TMA rows (P2) in conv3x3: parity, then A/B (and with the CTA-pair kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "conv3x3 or bf16 or dense_block or 1080p" 2>&1 | tail -15 > gpurun_out/r82_tests.log
cat gpurun_out/r82_tests.log
for cfg in "1 0" "0 0" "1 1"; do
  set -- $cfg
  SELFC_TC3_P2=$1 SELFC_TC3_PAIR=$2 timeout 600 python bench.py --steps 3 --warmup 3 --frames 28 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/r82_bench_p2$1_pair$2.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r82_bench_p2$1_pair$2.json"))
print("p2=$1 pair=$2", round(d["value"],1), "fps", {k:v["ms"] for k,v in d["roofline"]["classes"].items()}, d["clocks"]["sm_mhz"])
PY
done
