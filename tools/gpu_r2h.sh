#!/bin/bash
# round 2, call h: fused first two sweeps (box kernel MODE 4): parity + timing; then the whole GPU test-suite
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_patterns.py -m gpu -q -x -k "box_kernel or grid_hinted" 2>&1 | tail -5
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_patterns.py -m gpu -q -x -k "box_kernel and poisson and V and 11" > gpurun_out/r2h_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/r2h_memcheck.log
timeout 900 python tools/tune.py fuse_first=0 box_variant=11 > gpurun_out/r2h_tune.log 2>&1; echo "tune exit $?"
cut -c1-900 gpurun_out/r2h_tune.log
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r2h_pytest_gpu.log; cat gpurun_out/r2h_pytest_gpu.log
