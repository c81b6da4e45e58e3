#!/bin/bash
# round 2, call q (1 GPU): per-kernel times of the current build, ncu of the level-1 prolongation and restriction
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python tools/tune.py grid_transfers=0 > gpurun_out/r2q_tune.log 2>&1; echo "tune exit $?"
cut -c1-900 gpurun_out/r2q_tune.log
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:'gxp_kernel' --launch-skip 4 --launch-count 1 -f -o /tmp/r2q_p python tools/ncu_cycle.py > gpurun_out/r2q_ncu_p.log 2>&1
echo "ncu exit $?"
ncu -i /tmp/r2q_p.ncu-rep --page raw --csv > gpurun_out/r2q_ncu_gxp_raw.csv 2>/dev/null
ncu -i /tmp/r2q_p.ncu-rep --page source --csv --print-source sass > gpurun_out/r2q_ncu_gxp_source.csv 2>/dev/null
ls -la gpurun_out/r2q*
