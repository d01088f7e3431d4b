#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 2>gpurun_out/r90_err.log | grep '^{' | tail -1 > gpurun_out/r90_bench_4gpu.json; cut -c1-200 gpurun_out/r90_bench_4gpu.json; tail -3 gpurun_out/r90_err.log
python - <<PY
import json
d=json.load(open("gpurun_out/r90_bench_4gpu.json"))
print(d["n_gpus"], round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "u8", round(d["e2e"]["u8_frames"]["value"],1), d["clocks"])
PY
