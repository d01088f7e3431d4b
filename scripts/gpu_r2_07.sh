#!/bin/bash
mkdir -p gpurun_out
for e in 0 1 2 4 8 15; do echo "exp=$e"; SELFC_DBF_EXP=$e timeout 300 python scripts/dbg_fused2.py 2>&1 | tail -4 | head -1; done
