#!/bin/bash
# temporal kernel: output-frame-major passes with rolling accumulators + frame ring; GlobalAgg = pointwise GEMM + ga_mix
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r36_pytest.log
timeout 600 python bench.py --mode bf16 --frames 14 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r36_bench.json 2> gpurun_out/r36_err.txt
tail -3 gpurun_out/r36_err.txt
python - <<PY
import json
d=json.load(open('gpurun_out/r36_bench.json'))
print('value',d['value']); print(json.dumps(d['roofline']['classes']))
PY
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"temporal|ga_" -c 110 --csv --log-file gpurun_out/r36_launches.csv \
   python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r36_launches_bench.log 2>&1
