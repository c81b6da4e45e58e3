#!/bin/bash
# round 2, call d: box-stencil kernel with host-planned tile records - parity, timings of the variants, ncu of two of them
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_patterns.py -m gpu -q -x -k "box_kernel" 2>&1 | tail -5
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_patterns.py -m gpu -q -x -k "box_kernel and poisson and V and 1" > gpurun_out/r2d_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/r2d_memcheck.log
cd tools && timeout 900 python tune_box.py > ../gpurun_out/r2d_tune_box.log 2>&1; echo "tune exit $?"; cd ..
cut -c1-420 gpurun_out/r2d_tune_box.log
for v in 0 1; do
  MGB200_BOX_VARIANT=$v timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'box_kernel' -c 8 -f -o gpurun_out/r2d_box_v$v python tools/ncu_cycle.py > gpurun_out/r2d_ncu_v$v.log 2>&1
  echo "ncu v$v exit $?"
done
