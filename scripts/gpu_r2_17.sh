#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "^\[conv5|^\[1080p|^\[vid4|passed|failed|Error|assert" | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>$O/r2b_17_err.log | grep '^{' | tail -1 > $O/r2b_17_bench.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2b_17_bench.json"))
print(round(d["value"],1), "fps e2e", round(d["e2e"]["value"],1), d["clocks"], {k:(v["ms"],v["launches"]) for k,v in d["roofline"]["classes"].items()}, "frac", round(d["roofline"]["frac"],3))
PY
