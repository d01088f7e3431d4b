#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --height 576 --width 704 --frames 700 --steps 3 --warmup 3 --no-cpu-baseline 2>$O/r2b_14_err.log | grep '^{' | tail -1 > $O/r2b_14_vid4.json
timeout 300 python bench.py --height 2160 --width 3840 --frames 28 --steps 3 --warmup 3 --no-cpu-baseline 2>>$O/r2b_14_err.log | grep '^{' | tail -1 > $O/r2b_14_4k.json
python - <<PY
import json
for n in ("vid4","4k"):
    d=json.load(open("gpurun_out/r2b_14_%s.json"%n))
    print(n, d["metric"], round(d["value"],1), "fps e2e", round(d["e2e"]["value"],1), "u8", round(d["e2e"]["u8_frames"]["value"],1), d["clocks"], {k:v["ms"] for k,v in d["roofline"]["classes"].items()}, "frac", round(d["roofline"]["frac"],3), "launches", d["gpu_launches"])
PY
