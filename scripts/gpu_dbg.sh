#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q -s -k "test_bf16_mode_vs_oracle" 2>&1 | grep -E "^\[bf16|passed|failed|Error|assert" | head -20
