#!/bin/bash
# re-validation after container re-creation: parity tests, smoke, both conv3x3 variants, ncu launch list + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r20_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/r20_smoke.log
for v in 1 0; do
SELFC_TC_CONV2=$v timeout 600 python bench.py --mode bf16 --frames 14 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r20_bench_v$v.json 2> gpurun_out/r20_err_v$v.txt
python - <<PY
import json
d=json.load(open('gpurun_out/r20_bench_v$v.json'))
print('CONV2=$v value',d['value']); print(json.dumps(d['roofline']['classes']))
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file gpurun_out/r20_launches.csv \
   python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r20_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s 20 -c 4 -o gpurun_out/r20_prof_conv2 \
   python bench.py --mode bf16 --frames 7 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > gpurun_out/r20_prof_conv2.log 2>&1
ls -la gpurun_out
