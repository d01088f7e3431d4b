#!/bin/bash
# conv3x3 K-steps per pipeline stage with the pair + position-pair defaults (28-frame runs, interleaved twice)
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --mode bf16 --frames 28 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for rep in 1 2; do for k in 1 2 3 4; do
  SELFC_TC3_KPS=$k timeout 200 $B 2>/dev/null | grep '^{' | tail -1 > $O/r96_kps$k.json
  python - <<PY
import json
d=json.load(open("$O/r96_kps$k.json"))
c=d["roofline"]["classes"]
print("KPS=$k", round(d["value"],1), "fps", d["clocks"]["sm_mhz"], "conv3x3", c["conv3x3"]["ms"], "conv5", c["conv5_coupling"]["ms"])
PY
done; done
