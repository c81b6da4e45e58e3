#!/bin/bash
# round 2, call t (1 GPU): marching block kernel (cfg4), long-line box variants on 513^2 planes (Float64 slab shape and cfg5)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_patterns.py tests/test_gpu_parity.py tests/test_baseline_sizes.py -m gpu -x -q -k "round2 or block or cfg4" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/r2t_pytest.log
timeout 900 python tools/tune.py --cells 128 --levels 5 --nrhs 32 mrhs_march=0 > gpurun_out/r2t_tune_cfg4.log 2>&1; echo "tune cfg4 exit $?"
cut -c1-900 gpurun_out/r2t_tune_cfg4.log
timeout 900 python tools/tune.py --grid 512,512,64 --levels 6 box_variant=30 box_variant=31 box_variant=32 box_variant=33 box_variant=9 box_variant=11 > gpurun_out/r2t_tune_slab.log 2>&1; echo "tune slab exit $?"
cut -c1-420 gpurun_out/r2t_tune_slab.log
timeout 1500 python tools/tune.py --helmholtz --cells 512 --levels 7 box_variant_c=30 box_variant_c=31 box_variant_c=32 box_variant_c=9 box_variant_c=22 > gpurun_out/r2t_tune_cfg5.log 2>&1; echo "tune cfg5 exit $?"
cut -c1-420 gpurun_out/r2t_tune_cfg5.log
