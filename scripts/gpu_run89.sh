#!/bin/bash
# other configs of BASELINE.json on the round-2 build: Vid4 shape (bf16 + fp32 numerics mode), 4K
mkdir -p gpurun_out
timeout 600 python bench.py --height 576 --width 704 --frames 56 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r89_vid4_bf16.json
timeout 600 python bench.py --mode fp32 --height 576 --width 704 --frames 28 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r89_vid4_fp32.json
timeout 600 python bench.py --height 2160 --width 3840 --frames 14 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r89_4k_bf16.json
for f in vid4_bf16 vid4_fp32 4k_bf16; do python - <<PY
import json
d=json.load(open("gpurun_out/r89_$f.json"))
print("$f", round(d["value"],1), "fps e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, d["clocks"]["sm_mhz"])
PY
done
