#!/bin/bash
# Single-GPU round: the three inner-chain variants of the TMA kernel (bit-identity tests + bench each).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for v in 2 1 0; do
  MGB200_TMA_VARIANT=$v timeout 300 python -m pytest tests/test_patterns.py -m gpu -x -q > gpurun_out/pytest_patterns_v$v.log 2>&1; echo "pytest v$v exit $?"
  tail -1 gpurun_out/pytest_patterns_v$v.log
  MGB200_TMA_VARIANT=$v timeout 300 python bench.py --no-cpu > gpurun_out/bench_n1_v$v.json 2> gpurun_out/bench_n1_v$v.log; echo "bench v$v exit $?"
  cut -c1-200 gpurun_out/bench_n1_v$v.json
  grep "per-kernel" gpurun_out/bench_n1_v$v.log | cut -c1-800
done
