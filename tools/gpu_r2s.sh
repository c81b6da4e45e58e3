#!/bin/bash
# round 2, call s (1 GPU): device-side replaceMatrixInHierarchy tests, quad prolongation register budgets, cfg5 per-kernel times
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_replace_matrix.py tests/test_patterns.py -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/r2s_pytest.log
timeout 900 python tools/tune.py gxp_quad=2 gxp_quad=3 > gpurun_out/r2s_tune.log 2>&1; echo "tune exit $?"
cut -c1-700 gpurun_out/r2s_tune.log
timeout 1200 python tools/tune.py --helmholtz --cells 512 --levels 7 box_variant_c=1 > gpurun_out/r2s_tune_cfg5.log 2>&1; echo "tune cfg5 exit $?"
cut -c1-1200 gpurun_out/r2s_tune_cfg5.log
