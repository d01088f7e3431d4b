#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/dbg_fused.py 2>&1 | grep "schedule=0" | head -3 | cut -c1-400
