#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "global_agg or optimize_parameters or training_state or dense_block_bf16 or network_vs or bf16_mode_vs_oracle or model_wrapper" 2>&1 | tail -3
timeout 600 python bench.py --steps 3 --warmup 3 2>$O/r2b_16_err.log | grep '^{' | tail -1 > $O/r2b_16_bench.json
python - <<PY
import json
d=json.load(open("gpurun_out/r2b_16_bench.json"))
print(d["metric"], round(d["value"],1), "fps e2e", round(d["e2e"]["value"],1), "u8", round(d["e2e"]["u8_frames"]["value"],1), d["clocks"])
print({k:v["ms"] for k,v in d["roofline"]["classes"].items()}, "frac", round(d["roofline"]["frac"],3), "traffic", d["roofline"]["traffic"])
print("train", d["train"])
print("cpu", d["cpu_baseline"])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | grep '^{' | tail -1 | cut -c1-700
