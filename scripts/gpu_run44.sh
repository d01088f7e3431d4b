#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r44_pytest.log
for f in 1 0; do
SELFC_GMM_FUSED=$f timeout 600 python bench.py --mode bf16 --frames 14 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r44_bench_f$f.json 2> gpurun_out/r44_err.txt
tail -3 gpurun_out/r44_err.txt
python - <<PY
import json
d=json.load(open('gpurun_out/r44_bench_f$f.json'))
print('FUSED=$f value',d['value']); print(json.dumps(d['roofline']['classes']))
PY
done
