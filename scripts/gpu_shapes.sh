#!/bin/bash
# other shapes of BASELINE.json's configs with the current build:  gpurun --timeout 1500 -- 'bash scripts/gpu_shapes.sh r2c'
TAG=${1:-run}
O=gpurun_out
mkdir -p $O
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-train"
timeout 300 $B --frames 28 --no-e2e 2>/dev/null | grep '^{' | tail -1 > $O/${TAG}_bench_bf16_28frames.json
for g in 1 4 10; do timeout 300 $B --height 576 --width 704 --frames 700 --gops-per-launch $g 2>/dev/null | grep '^{' | tail -1 > $O/${TAG}_bench_vid4_gpl$g.json; done
timeout 300 $B --height 2160 --width 3840 --frames 28 2>/dev/null | grep '^{' | tail -1 > $O/${TAG}_bench_4k_bf16.json
timeout 300 $B --mode fp32 --height 576 --width 704 --frames 28 --no-e2e 2>/dev/null | grep '^{' | tail -1 > $O/${TAG}_bench_vid4_fp32.json
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/${TAG}_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], d.get("metric"), round(d["value"], 1), d.get("clocks", {}).get("sm_mhz"), "frac", round(d["roofline"]["frac"], 3) if d.get("roofline") else None)
    except Exception as e:
        print(f, "unreadable", e)
PY
