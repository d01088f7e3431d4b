#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dbg_fused.py > gpurun_out/r2b_06_dbg_fused.txt 2>&1; echo rc=$?; cat gpurun_out/r2b_06_dbg_fused.txt | tail -20
