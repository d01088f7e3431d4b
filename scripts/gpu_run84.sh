#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r84_pytest.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r84_bench.json 2> gpurun_out/r84_bench.err; tail -2 gpurun_out/r84_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/r84_bench.json"))
print(round(d["value"],1), "fps e2e", round(d["e2e"]["value"],1), {k:v["ms"] for k,v in d["roofline"]["classes"].items()}, d["clocks"], d["roofline"]["frac"])
PY
