#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --frames 14 --steps 2 --warmup 3 > gpurun_out/r67_b.json 2> gpurun_out/r67_err.txt; tail -2 gpurun_out/r67_err.txt; cut -c1-200 gpurun_out/r67_b.json
timeout 600 python bench.py --workload train --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-200
