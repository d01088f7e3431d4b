#!/bin/bash
# full -m gpu suite + smoke with the fused dense-block kernel as the default path
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2b_10_pytest.log 2>&1; echo rc=$?
grep -E "^\[|passed|failed|Error|error|assert" gpurun_out/r2b_10_pytest.log | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_10_smoke.log 2>&1; echo smoke rc=$?; tail -5 gpurun_out/r2b_10_smoke.log
