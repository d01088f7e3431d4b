#!/bin/bash
for kc in 64 32 16; do
echo "=== KC=$kc"
SELFC_TC_KC=$kc timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dense_block_bf16 or global_agg_bf16 or smoke" 2>&1 | tail -2
SELFC_TC_KC=$kc timeout 600 python bench.py --mode bf16 --frames 14 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r10_bench_kc$kc.json 2> gpurun_out/r10_err.txt; python - <<PY
import json
d=json.load(open('gpurun_out/r10_bench_kc$kc.json'))
print('value',d['value'])
print(json.dumps(d['roofline']['classes']))
PY
SELFC_TC_KC=$kc timeout 300 python scripts/dbg_temporal.py | head -3
done
