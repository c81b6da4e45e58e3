#!/bin/bash
# round 2, call k: tolerance trial of the Krylov / coarsest-GMRES parity tests, complex variants of the box kernel, ncu of the prolongation
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -12
timeout 900 python tools/tune.py --helmholtz --cells 256 --levels 6 box_variant_c=0 box_variant_c=20 box_variant_c=21 box_variant_c=22 box_variant_c=9 box_variant_c=3 > gpurun_out/r2k_tune_c.log 2>&1; echo "tune exit $?"
cut -c1-600 gpurun_out/r2k_tune_c.log
MGB200_BOX_VARIANT=1 timeout 600 ncu --set full --clock-control none --profile-from-start off \
      -k regex:'gxp_kernel' -c 2 -f -o /tmp/r2k python tools/ncu_cycle.py > gpurun_out/r2k_ncu.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/r2k.ncu-rep --page raw --csv > gpurun_out/r2k_ncu_gxp_raw.csv 2>/dev/null
