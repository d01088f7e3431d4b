#!/bin/bash
mkdir -p gpurun_out
# memcheck over the small-shape component tests (OOB / misaligned accesses in any kernel)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "global_agg or gmm_sample or philox or conv3x3_tc or dense_block_bf16 or network_vs_golden or bf16_mode" 2>&1 | tail -25 > gpurun_out/r48_memcheck.txt
tail -12 gpurun_out/r48_memcheck.txt
