#!/bin/bash
# session-4 build on the other BASELINE configs: Vid4 shape (bf16), 4K, training step
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python bench.py --height 576 --width 704 --frames 56 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 > $O/r95_vid4_bf16.json
timeout 200 python bench.py --height 2160 --width 3840 --frames 14 --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 > $O/r95_4k_bf16.json
timeout 200 python bench.py --workload train --steps 5 --warmup 3 2>/dev/null | grep '^{' | tail -1 > $O/r95_train_1gpu.json
for f in vid4_bf16 4k_bf16 train_1gpu; do python - <<PY
import json
d=json.load(open("$O/r95_$f.json"))
print("$f", round(d["value"],2), d["unit"], "e2e", round(d["e2e"]["value"],1) if d.get("e2e") else None, d["clocks"]["sm_mhz"])
PY
done
