#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 2>gpurun_out/r2c_2gpu_err.log | grep '^{' | tail -1 > gpurun_out/r2c_bench_2gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2c_bench_2gpu.json"))
print(d["n_gpus"], round(d["value"],1), "fps e2e", round(d["e2e"]["value"],1), "fp32 host frames", round(d["e2e"]["fp32_frames"]["value"],1), d["clocks"], "train", {k:d["train"][k] for k in ("septuplets_per_s","ms_per_step","allreduce_ms","n_gpus")})
PY
tail -3 gpurun_out/r2c_2gpu_err.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 2>/dev/null | grep '^{' | cut -c1-120
