#!/bin/bash
# Final single-GPU round of the session: whole parity suite, bench line, Float64 / Float32 / mixed comparison,
# ncu launch list of the bench command, full capture of the dictionary kernels of one cycle.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench exit $?"
cut -c1-300 gpurun_out/bench_n1.json
timeout 300 python tools/bench_precision.py > gpurun_out/bench_precision.json 2> gpurun_out/bench_precision.log; echo "bench_precision exit $?"
cut -c1-1500 gpurun_out/bench_precision.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_ncu_launches.log 2>&1
echo "ncu launches exit $? rows $(grep -c gpu__time_duration gpurun_out/launches_bench_steps2.csv)"
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'pat_tma|pat_kernel' -c 12 \
    -f -o gpurun_out/pat_full_r01i python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_ncu_full.log 2>&1
echo "ncu full exit $?"
