#!/bin/bash
mkdir -p gpurun_out
for g in 1 3 5; do
timeout 600 python bench.py --mode bf16 --frames 105 --steps 2 --warmup 2 --gops-per-launch $g --no-cpu-baseline --no-e2e > gpurun_out/r45_bench_g$g.json 2> gpurun_out/r45_err.txt
tail -3 gpurun_out/r45_err.txt
python - <<PY
import json
d=json.load(open('gpurun_out/r45_bench_g$g.json'))
print('GPL=$g value',d['value'], d['clocks'])
PY
done
