#!/bin/bash
# round 2, call b: box-stencil kernel - parity on the GPU (bit-identity against the dictionary walk), memcheck of one
# case, then timings of its variants on cfg2
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_patterns.py -m gpu -q -x -k "box_kernel" 2>&1 | tail -15
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_patterns.py -m gpu -q -x -k "box_kernel and poisson and V and 3" > gpurun_out/r2b_memcheck.log 2>&1; echo "memcheck exit $?"; tail -5 gpurun_out/r2b_memcheck.log
cd tools && timeout 900 python tune_box.py > ../gpurun_out/r2b_tune_box.log 2>&1; echo "tune exit $?"; cd ..
cut -c1-600 gpurun_out/r2b_tune_box.log
