#!/bin/bash
# mbarrier try_wait with a suspend-time hint: sustained (power-capped) default bench + a quick parity subset
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "bf16 or 1080p or variants" 2>&1 | tail -3 | tee gpurun_out/r85_tests.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r85_bench.json 2> gpurun_out/r85_bench.err; tail -2 gpurun_out/r85_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r85_bench.json"))
print(round(d["value"],1), "fps e2e", round(d["e2e"]["value"],1), {k:v["ms"] for k,v in d["roofline"]["classes"].items()}, d["clocks"], d["roofline"]["frac"])
PY
