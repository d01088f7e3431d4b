#!/bin/bash
# end-of-round validation: all parity tests, smoke, default bench + reference arm, ncu launch list of one GOP, full captures
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r64_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/r64_smoke.log
timeout 900 python bench.py > gpurun_out/r64_bench.json 2> gpurun_out/r64_bench.err; tail -2 gpurun_out/r64_bench.err; cut -c1-300 gpurun_out/r64_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r64_ref.json 2>/dev/null
B="python bench.py --mode bf16 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 440 -c 560 --csv --log-file gpurun_out/r64_launches.csv $B > gpurun_out/r64_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc3 -s 16 -c 8 -o gpurun_out/r64_conv3 $B > gpurun_out/r64_n1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:temporal_tc_kernel -s 16 -c 6 -o gpurun_out/r64_temporal $B > gpurun_out/r64_n2.log 2>&1
ls -la gpurun_out | grep r64
