#!/bin/bash
for shp in "1 2 36 56" "1 3 148 484" "1 1 1080 1920"; do echo "== $shp"; timeout 300 python scripts/dbg_f5.py $shp 2>&1 | tail -7 | cut -c1-300; done
