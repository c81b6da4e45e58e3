#!/bin/bash
# round 2, call e: new grid-hinted transfer kernels (parity + timing), ncu of the box kernel (raw csv, the .ncu-rep stays on the box)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_patterns.py -m gpu -q -x -k "grid_hinted" 2>&1 | tail -5
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_patterns.py -m gpu -q -x -k "grid_hinted and 40" > gpurun_out/r2e_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/r2e_memcheck.log
timeout 900 python tools/tune.py grid_transfers=0 box=0,grid_transfers=0 box_variant=0 box_variant=3 > gpurun_out/r2e_tune.log 2>&1; echo "tune exit $?"
cut -c1-1200 gpurun_out/r2e_tune.log
for v in 0 1; do
  MGB200_BOX_VARIANT=$v timeout 600 ncu --set full --clock-control none --profile-from-start off \
      -k regex:'box_kernel|gxp_kernel|gxr_kernel' -c 10 -f -o /tmp/r2e_v$v python tools/ncu_cycle.py > gpurun_out/r2e_ncu_v$v.log 2>&1
  echo "ncu v$v exit $?"
  ncu -i /tmp/r2e_v$v.ncu-rep --page raw --csv > gpurun_out/r2e_ncu_v${v}_raw.csv 2>/dev/null
  ncu -i /tmp/r2e_v$v.ncu-rep --page details > gpurun_out/r2e_ncu_v${v}_details.txt 2>/dev/null
done
ls -la gpurun_out | tail -12
