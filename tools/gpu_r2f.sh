#!/bin/bash
# round 2, call f: direct (no staging) box kernel variants 8-10 and the restriction fast path: parity, memcheck, timing, ncu
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_patterns.py -m gpu -q -x -k "box_kernel or grid_hinted" 2>&1 | tail -5
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_patterns.py -m gpu -q -x -k "box_kernel and poisson and V and 9" > gpurun_out/r2f_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/r2f_memcheck.log
timeout 900 python tools/tune.py box_variant=1 box_variant=8 box_variant=9 box_variant=10 > gpurun_out/r2f_tune.log 2>&1; echo "tune exit $?"
cut -c1-700 gpurun_out/r2f_tune.log
for v in 8 9; do
  MGB200_BOX_VARIANT=$v timeout 600 ncu --set full --clock-control none --profile-from-start off \
      -k regex:'box_direct_kernel|gxp_kernel|gxr_kernel' -c 14 -f -o /tmp/r2f_v$v python tools/ncu_cycle.py > gpurun_out/r2f_ncu_v$v.log 2>&1
  echo "ncu v$v exit $?"
  ncu -i /tmp/r2f_v$v.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_v${v}_raw.csv 2>/dev/null
done
ls -la gpurun_out | tail -6
