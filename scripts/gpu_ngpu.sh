#!/bin/bash
# N-GPU bench under torchrun, as the driver launches it:  gpurun --gpus N -- 'bash scripts/gpu_ngpu.sh N <tag>'
N=${1:-2}; TAG=${2:-run}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 2>gpurun_out/${TAG}_${N}gpu_err.log | grep '^{' | tail -1 > gpurun_out/${TAG}_bench_${N}gpu.json
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_${N}gpu.json"))
print(d["n_gpus"], "GPUs:", round(d["value"],1), "fps, e2e", round(d["e2e"]["value"],1), "(fp32 host frames", round(d["e2e"]["fp32_frames"]["value"],1), ")", d["clocks"], "train", {k:round(d["train"][k],3) for k in ("septuplets_per_s","ms_per_step","allreduce_ms")})
PY
tail -2 gpurun_out/${TAG}_${N}gpu_err.log | cut -c1-200
