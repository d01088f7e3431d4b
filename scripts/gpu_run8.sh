#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bf16 or drop_in or tc" 2>&1 | tail -8
timeout 600 python bench.py --mode bf16 --frames 14 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r11_bench_bf16.json 2> gpurun_out/r11_bench_bf16.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r11_bench_bf16.json'))
print('value',d['value'],'e2e',d['e2e']['value'])
print(json.dumps(d['roofline']['classes']))
PY
tail -3 gpurun_out/r11_bench_bf16.err
