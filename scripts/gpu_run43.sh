#!/bin/bash
# checkpoint after the temporal-kernel / GlobalAgg work: parity, default bench, reference arm, ncu launch list + full captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r43_pytest.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/r43_smoke.log
timeout 900 python bench.py > gpurun_out/r43_bench.json 2> gpurun_out/r43_bench.err; tail -2 gpurun_out/r43_bench.err; cut -c1-600 gpurun_out/r43_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r43_ref.json 2>/dev/null
B="python bench.py --mode bf16 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 440 -c 700 --csv --log-file gpurun_out/r43_launches.csv $B > gpurun_out/r43_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc3 -s 20 -c 4 -o gpurun_out/r43_conv3 $B > gpurun_out/r43_n1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"temporal_tc_kernel<3>" -s 24 -c 3 -o gpurun_out/r43_temporal3 $B > gpurun_out/r43_n2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"temporal_tc_kernel<1>|ga_mix|ga_stat|ga_weights|gmm_sample" -s 0 -c 9 -o gpurun_out/r43_stp $B > gpurun_out/r43_n3.log 2>&1
ls -la gpurun_out | grep r43
