#!/bin/bash
# slab-planar dense buffers + conv3x3 v3 (pixels in M, kx stacked in N): parity, then per-class timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r21_pytest.log
for kps in 2 4 1; do
SELFC_TC3_KPS=$kps timeout 600 python bench.py --mode bf16 --frames 14 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r21_bench_k$kps.json 2> gpurun_out/r21_err_k$kps.txt
tail -3 gpurun_out/r21_err_k$kps.txt
python - <<PY
import json
d=json.load(open('gpurun_out/r21_bench_k$kps.json'))
print('KPS=$kps value',d['value']); print(json.dumps(d['roofline']['classes']))
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv3x3|temporal" -c 330 --csv --log-file gpurun_out/r21_launches.csv \
   python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r21_launches_bench.log 2>&1
