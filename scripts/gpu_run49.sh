#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for dgh in 1 0; do
SELFC_DUAL_GH=$dgh timeout 600 python bench.py --mode bf16 --frames 28 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r49_bench_$dgh.json 2> gpurun_out/r49_err.txt
python - <<PY
import json
d=json.load(open('gpurun_out/r49_bench_$dgh.json'))
print('DUAL=$dgh value %.1f'%d['value'], 'conv3x3 %.2f ms'%d['roofline']['classes']['conv3x3']['ms'], d['clocks']['sm_mhz'])
PY
done
done
for i in $(seq 1 15); do timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gmm_sample" 2>&1 | tail -1; done
