#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "optimize_parameters or optimizer_step or train_step" 2>&1 | tail -30
