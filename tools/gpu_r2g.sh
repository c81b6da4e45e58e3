#!/bin/bash
# round 2, call g: big-tile box variants 11-13, prolongation with 4 lines per thread: parity, timing, ncu of the transfers
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_patterns.py -m gpu -q -x -k "box_kernel or grid_hinted" 2>&1 | tail -5
timeout 900 python tools/tune.py box_variant=11 box_variant=12 box_variant=13 box_variant=1,box_variant27=3 box_variant=11,box_variant27=1 > gpurun_out/r2g_tune.log 2>&1; echo "tune exit $?"
cut -c1-700 gpurun_out/r2g_tune.log
MGB200_BOX_VARIANT=11 timeout 600 ncu --set full --clock-control none --profile-from-start off \
      -k regex:'box_kernel|gxp_kernel|gxr_kernel|diag_scale' -c 12 -f -o /tmp/r2g python tools/ncu_cycle.py > gpurun_out/r2g_ncu.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/r2g.ncu-rep --page raw --csv > gpurun_out/r2g_ncu_raw.csv 2>/dev/null
ls -la gpurun_out | tail -4
