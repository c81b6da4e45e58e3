#!/bin/bash
# Single-GPU round: single-precision tests (Float32 now takes the TMA kernel too), dictionary bit-identity tests,
# and the Float64 / Float32 / mixed-precision comparison on cfg2.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_single_precision.py tests/test_patterns.py -m gpu -q > gpurun_out/pytest_single.log 2>&1; echo "pytest single+patterns exit $?"
tail -5 gpurun_out/pytest_single.log | cut -c1-220
timeout 600 python tools/bench_precision.py > gpurun_out/bench_precision.json 2> gpurun_out/bench_precision.log; echo "bench_precision exit $?"
cut -c1-1800 gpurun_out/bench_precision.json
tail -3 gpurun_out/bench_precision.log | cut -c1-300
