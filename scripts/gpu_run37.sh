#!/bin/bash
timeout 300 python scripts/dbg_temporal.py 2>&1 | tail -14
