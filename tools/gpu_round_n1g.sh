#!/bin/bash
# Single-GPU round: single-precision tests first (new), then the whole parity suite.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_single_precision.py -m gpu -q > gpurun_out/pytest_single.log 2>&1; echo "pytest single exit $?"
tail -40 gpurun_out/pytest_single.log | cut -c1-220
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_single_precision.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
