#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r51_pytest.log
for f in 1 0; do
SELFC_FUSE_HG=$f timeout 600 python bench.py --mode bf16 --frames 28 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r51_bench_$f.json 2> gpurun_out/r51_err.txt
tail -2 gpurun_out/r51_err.txt
python - <<PY
import json
d=json.load(open('gpurun_out/r51_bench_$f.json'))
print('FUSE_HG=$f value %.1f'%d['value'], json.dumps(d['roofline']['classes']['conv5_coupling']), d['clocks']['sm_mhz'])
PY
done
