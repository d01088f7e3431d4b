#!/bin/bash
# round 2, run 4: fused vs layer-by-layer dense blocks, 1080p bf16, 28-frame runs + one 100-frame run
mkdir -p gpurun_out
for f in 1 0; do
  SELFC_DB_FUSED=$f timeout 300 python bench.py --frames 28 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/r2b_04_err$f.log | grep '^{' | tail -1 > gpurun_out/r2b_04_fused$f.json
  python - <<PY
import json
d=json.load(open("gpurun_out/r2b_04_fused$f.json"))
print("fused=$f", round(d["value"],1), "fps", d["clocks"], {k:(v["ms"],v["launches"]) for k,v in d["roofline"]["classes"].items()}, "frac", round(d["roofline"]["frac"],3))
PY
done
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/r2b_04_err.log | grep '^{' | tail -1 > gpurun_out/r2b_04_bench100.json
python -c "
import json
d=json.load(open('gpurun_out/r2b_04_bench100.json'))
print('100 frames', round(d['value'],1), 'fps e2e', round(d['e2e']['value'],1), d['clocks'], 'frac', round(d['roofline']['frac'],3))"
