#!/bin/bash
# 1 GPU: automatic box variants, marching block kernel with prefetch, block transfer kernels
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests/test_patterns.py tests/test_replace_matrix.py tests/test_gpu_parity.py tests/test_baseline_sizes.py -m gpu -x -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/r2u_pytest.log
timeout 900 python tools/tune.py box_variant=11 > gpurun_out/r2u_tune.log 2>&1; echo "tune exit $?"
cut -c1-700 gpurun_out/r2u_tune.log
timeout 900 python tools/tune.py --cells 128 --levels 5 --nrhs 32 grid_transfers=0 > gpurun_out/r2u_tune_cfg4.log 2>&1; echo "tune cfg4 exit $?"
cut -c1-900 gpurun_out/r2u_tune_cfg4.log
timeout 900 python tools/tune.py --grid 512,512,64 --levels 6 > gpurun_out/r2u_tune_slab.log 2>&1; echo "tune slab exit $?"
cut -c1-600 gpurun_out/r2u_tune_slab.log
timeout 1500 python tools/tune.py --helmholtz --cells 512 --levels 7 > gpurun_out/r2u_tune_cfg5.log 2>&1; echo "tune cfg5 exit $?"
cut -c1-700 gpurun_out/r2u_tune_cfg5.log
