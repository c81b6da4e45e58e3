#!/bin/bash
# Single-GPU round: GPU parity tests, ncu launch list of the bench command, full capture of the dictionary kernels
# of one cycle (levels 1-3).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_ncu_launches.log 2>&1
echo "ncu launches exit $? rows $(grep -c gpu__time_duration gpurun_out/launches_bench_steps2.csv)"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'pat_tma|pat_kernel' -c 16 \
    -f -o gpurun_out/pat_full python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out | head -30
