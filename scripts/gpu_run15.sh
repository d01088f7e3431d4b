#!/bin/bash
for d in 0 1 2 3; do
SELFC_TC2_DBG=$d timeout 600 python bench.py --mode bf16 --frames 14 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r15_bench_d$d.json 2> gpurun_out/r15_err.txt; python - <<PY
import json
d=json.load(open('gpurun_out/r15_bench_d$d.json'))
print('DBG=$d', json.dumps(d['roofline']['classes']['conv3x3']))
PY
done
