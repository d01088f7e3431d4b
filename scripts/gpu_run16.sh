#!/bin/bash
for kc in 32 64 16; do
echo "== KC=$kc"
SELFC_TC2_KC=$kc timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "conv3x3_tc or bf16_mode" 2>&1 | tail -3
for d in 0 3; do
SELFC_TC2_KC=$kc SELFC_TC2_DBG=$d timeout 600 python bench.py --mode bf16 --frames 14 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r16_b.json 2> gpurun_out/r16_err.txt; python - <<PY
import json
d=json.load(open('gpurun_out/r16_b.json'))
print('KC=$kc DBG=$d value', round(d['value'],1), json.dumps(d['roofline']['classes']['conv3x3']))
PY
done
done
