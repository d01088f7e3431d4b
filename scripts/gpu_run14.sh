#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv3x3_tc" 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16 or drop_in" 2>&1 | tail -4
for v in 1 0; do
SELFC_TC_CONV2=$v timeout 600 python bench.py --mode bf16 --frames 28 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r14_bench_v$v.json 2> gpurun_out/r14_err.txt; python - <<PY
import json
d=json.load(open('gpurun_out/r14_bench_v$v.json'))
print('CONV2=$v value',d['value'])
print(json.dumps(d['roofline']['classes']['conv3x3']))
PY
done
tail -3 gpurun_out/r14_err.txt
