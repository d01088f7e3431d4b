#!/bin/bash
# sampler forms A/B (0 thread-per-pixel, 1 warp-split, 2 warp-split per component), in-stream class times, 28-frame runs
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "gmm_sample or global_agg" > $O/r93_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/r93_pytest_gpu.log
B="python bench.py --mode bf16 --frames 28 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for f in 0 1 2 0 1 2; do
  SELFC_GMM_SPLIT=$f timeout 200 $B 2>/dev/null | grep '^{' | tail -1 > $O/r93_ab_$f.json
  python - <<PY
import json
d=json.load(open("$O/r93_ab_$f.json"))
c=d["roofline"]["classes"]
print("SELFC_GMM_SPLIT=$f", round(d["value"],1), "fps", d["clocks"]["sm_mhz"], "sampler", c["sampler"]["ms"], "ga", c["global_agg"]["ms"])
PY
done
B1="python bench.py --mode bf16 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
SELFC_GMM_SPLIT=2 timeout 120 ncu --set full --clock-control none --import-source on -k regex:gmm_sample_planar -s 1 -c 1 -o $O/r93_sampler_perk $B1 > $O/r93_n1.log 2>&1
ls -la $O | grep r93
