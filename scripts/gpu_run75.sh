#!/bin/bash
# round-2 validation: all parity tests, smoke, default bench + reference arm
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r75_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/r75_smoke.log
timeout 900 python bench.py > gpurun_out/r75_bench.json 2> gpurun_out/r75_bench.err; tail -2 gpurun_out/r75_bench.err; cut -c1-400 gpurun_out/r75_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r75_ref.json; cut -c1-300 gpurun_out/r75_ref.json
timeout 600 python bench.py --workload train --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r75_train.json; cut -c1-300 gpurun_out/r75_train.json
