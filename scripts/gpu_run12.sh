#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16 or drop_in or tc" 2>&1 | tail -4
for nopdl in 0 1; do
SELFC_NO_PDL=$nopdl timeout 600 python bench.py --mode bf16 --frames 28 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/r12_bench_nopdl$nopdl.json 2> gpurun_out/r12_err.txt; python - <<PY
import json
d=json.load(open('gpurun_out/r12_bench_nopdl$nopdl.json'))
print('NO_PDL=$nopdl value',d['value'])
print(json.dumps(d['roofline']['classes']))
PY
done
tail -3 gpurun_out/r12_err.txt
