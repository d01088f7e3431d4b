#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r13_bench.json 2> gpurun_out/r13_bench.err; tail -c 2600 gpurun_out/r13_bench.json; tail -3 gpurun_out/r13_bench.err
