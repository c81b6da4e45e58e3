#!/bin/bash
# Single-GPU round: parity suite + bench of the unrolled-chain TMA kernel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.log; echo "bench exit $?"
cut -c1-300 gpurun_out/bench_n1.json
grep "per-kernel" gpurun_out/bench_n1.log | cut -c1-1600
