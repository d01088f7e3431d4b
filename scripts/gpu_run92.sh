#!/bin/bash
# round-1 session 4: warp-split planar sampler + staged-weights ga_weights: parity, A/B, default bench, launch list, sampler capture
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $O/r92_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/r92_pytest_gpu.log
B="python bench.py --mode bf16 --frames 28 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
for cfg in "SELFC_GMM_SPLIT=0 SELFC_GA_WEIGHTS_V1=1" "SELFC_GMM_SPLIT=1 SELFC_GA_WEIGHTS_V1=1" "SELFC_GMM_SPLIT=1 SELFC_GA_WEIGHTS_V1=0"; do
  tag=$(echo $cfg | tr -d ' =A-Z_')
  env $cfg timeout 200 $B 2>/dev/null | grep '^{' | tail -1 > $O/r92_ab_$tag.json
  python - <<PY
import json
d=json.load(open("$O/r92_ab_$tag.json"))
c=d["roofline"]["classes"]
print("$cfg", round(d["value"],1), "fps", d["clocks"]["sm_mhz"], "sampler", c["sampler"]["ms"], "ga", c["global_agg"]["ms"])
PY
done
timeout 300 python bench.py 2>$O/r92_bench_err.log | grep '^{' | tail -1 > $O/r92_bench_bf16_100frames.json; cut -c1-160 $O/r92_bench_bf16_100frames.json
B1="python bench.py --mode bf16 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 440 -c 560 --csv --log-file $O/r92_launches.csv $B1 > $O/r92_launches_bench.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:gmm_sample_planar -s 1 -c 1 -o $O/r92_sampler $B1 > $O/r92_n1.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r92_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/r92_smoke.log
ls -la $O | grep r92
