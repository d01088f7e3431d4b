#!/bin/bash
# round-1 session 4 validation: full GPU parity suite, default bench, reference arm, launch list of one GOP, captures of the
# sampler / ga_stat / ga_weights, smoke
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > $O/r94_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $O/r94_pytest_gpu.log
timeout 300 python bench.py 2>$O/r94_bench_err.log | grep '^{' | tail -1 > $O/r94_bench_bf16_100frames.json; cut -c1-160 $O/r94_bench_bf16_100frames.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | grep '^{' | tail -1 > $O/r94_bench_reference_arm.json; cut -c1-200 $O/r94_bench_reference_arm.json
B28="python bench.py --mode bf16 --frames 28 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 200 $B28 2>/dev/null | grep '^{' | tail -1 > $O/r94_bench_bf16_28frames.json; cut -c1-160 $O/r94_bench_bf16_28frames.json
B1="python bench.py --mode bf16 --frames 7 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 440 -c 560 --csv --log-file $O/r94_launches.csv $B1 > $O/r94_launches_bench.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:"gmm_sample_planar|ga_stat|ga_weights" -s 3 -c 3 -o $O/r94_stp_small $B1 > $O/r94_n1.log 2>&1
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r94_smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/r94_smoke.log
ls -la $O | grep r94
