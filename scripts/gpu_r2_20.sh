#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused_dense_block or variants" > $O/r2c_t.log 2>&1; grep -E "^E  |passed|failed|^\[" $O/r2c_t.log | head -20
timeout 900 compute-sanitizer --tool racecheck --kernel-name kns=dense_fused --error-exitcode 3 python scripts/sanitize_small.py > $O/r2c_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -c "Race reported" $O/r2c_racecheck.log; grep "Race reported" $O/r2c_racecheck.log | grep -v "0xfffffffffffff" | head -5; tail -3 $O/r2c_racecheck.log
