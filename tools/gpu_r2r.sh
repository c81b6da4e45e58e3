#!/bin/bash
# round 2, call r (1 GPU): quad prolongation + 16-byte-load restriction: bit-identity tests, per-kernel times, ncu
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_patterns.py -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r2r_pytest.log
timeout 900 python tools/tune.py gxp_quad=0 gxr_vec=0 > gpurun_out/r2r_tune.log 2>&1; echo "tune exit $?"
cut -c1-1200 gpurun_out/r2r_tune.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'gxp_quad_kernel|gxr_kernel' --launch-count 2 -f -o /tmp/r2r python tools/ncu_cycle.py > gpurun_out/r2r_ncu.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/r2r.ncu-rep --page raw --csv > gpurun_out/r2r_ncu_raw.csv 2>/dev/null
ls -la gpurun_out/r2r*
