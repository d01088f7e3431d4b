#!/bin/bash
# checkpoint: full default bench (100 frames, e2e, cpu baseline) + reference arm, on main (slab layout, conv v3, 4-normal Philox)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r35_bench.json 2> gpurun_out/r35_bench.err; tail -2 gpurun_out/r35_bench.err; cat gpurun_out/r35_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r35_ref.json 2>/dev/null; cat gpurun_out/r35_ref.json
