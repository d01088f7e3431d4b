#!/bin/bash
mkdir -p gpurun_out
for cfg in "1 0" "0 0"; do
  set -- $cfg
  echo "== P2=$1 PAIR=$2" >> gpurun_out/r83_dbg.log
  SELFC_TC3_P2=$1 SELFC_TC3_PAIR=$2 python scripts/dbg_conv3.py >> gpurun_out/r83_dbg.log 2>&1
done
grep "==\|dual=0 nks=9\|dual=0 nks=3\|dual=1 nks=1" gpurun_out/r83_dbg.log | head -40
