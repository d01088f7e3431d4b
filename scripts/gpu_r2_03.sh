#!/bin/bash
# round 2, run 3: fused dense-block kernel -- first contact
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s -k "fused_dense_block" > gpurun_out/r2b_03_pytest.log 2>&1; echo rc=$?
grep -E "^\[|passed|failed|Error|error|assert|differs" gpurun_out/r2b_03_pytest.log | head -40
tail -5 gpurun_out/r2b_03_pytest.log
