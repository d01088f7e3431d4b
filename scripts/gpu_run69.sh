#!/bin/bash
# conv3x3 barrier-wait counters (timing build), then restore the normal build
mkdir -p gpurun_out
python scripts/dbg_conv3.py > gpurun_out/r69_dbg.log 2>&1
cat gpurun_out/r69_dbg.log | tail -20
