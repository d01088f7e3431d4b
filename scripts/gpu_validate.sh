#!/bin/bash
# Everything the driver runs at round end, on one B200:  gpurun --timeout 2400 -- 'bash scripts/gpu_validate.sh r2c'
# -> gpurun_out/<tag>_pytest_gpu.log, <tag>_smoke.log, <tag>_bench_bf16_100frames.json, <tag>_bench_reference_arm.json
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -s > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 $O/${TAG}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $O/${TAG}_smoke.log
timeout 600 python bench.py 2>$O/${TAG}_bench_err.log | grep '^{' | tail -1 > $O/${TAG}_bench_bf16_100frames.json; cut -c1-200 $O/${TAG}_bench_bf16_100frames.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | grep '^{' | tail -1 > $O/${TAG}_bench_reference_arm.json; cut -c1-160 $O/${TAG}_bench_reference_arm.json
