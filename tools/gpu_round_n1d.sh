#!/bin/bash
# Single-GPU round: parity suite, bench with the merged TMA windows (default) and with the round-1 windows (gap 32).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench exit $?"
cut -c1-300 gpurun_out/bench_n1.json
grep "per-kernel" gpurun_out/bench_n1.log | cut -c1-1200
MGB200_TMA_GAP=32 timeout 600 python bench.py --no-cpu > gpurun_out/bench_n1_gap32.json 2> gpurun_out/bench_n1_gap32.log; echo "bench gap32 exit $?"
cut -c1-300 gpurun_out/bench_n1_gap32.json
grep "per-kernel" gpurun_out/bench_n1_gap32.log | cut -c1-1200
