#!/bin/bash
# round 2, run 2: the new benchmark-size parity tests against the oracle
mkdir -p gpurun_out
nproc
timeout 900 python -m pytest tests -m gpu -x -q -s -k "1080p_gop or 1080p_two or vid4_shape_bf16 or test_bf16_mode_vs_oracle" > gpurun_out/r2b_02_pytest.log 2>&1; echo rc=$?
grep -E "^\[|passed|failed|Error|assert" gpurun_out/r2b_02_pytest.log | head -40
