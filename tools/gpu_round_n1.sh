#!/bin/bash
# One single-GPU round on the B200 box: GPU parity tests, the bench line, the ncu launch list of the same command and
# one full capture of the dominant kernel.  Outputs land in gpurun_out/.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench exit $?"
cat gpurun_out/bench_n1.json | head -c 3000
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pat_tma -c 6 \
    -f -o gpurun_out/pat_tma_full python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
