#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --mode bf16 --frames 28 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r65_bench.json 2> gpurun_out/r65_err.txt
tail -2 gpurun_out/r65_err.txt
python - <<PY
import json
d=json.load(open('gpurun_out/r65_bench.json'))
print('value %.1f'%d['value'], json.dumps(d['roofline']['classes']), d['clocks']['sm_mhz'])
PY
