#!/bin/bash
# round 2, call i: blocked LU + inversion (parity), prolongation lines per thread, then cfg3 timing
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "coarsest or sa_amg or solveMG_per_cycle or block_solveMG" 2>&1 | tail -5
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "coarsest_dense_lu" > gpurun_out/r2i_memcheck.log 2>&1; echo "memcheck exit $?"; tail -3 gpurun_out/r2i_memcheck.log
timeout 900 python tools/tune.py gxp_lines=2 gxp_lines=4 > gpurun_out/r2i_tune.log 2>&1; echo "tune exit $?"
cut -c1-420 gpurun_out/r2i_tune.log
timeout 1200 python tools/bench_configs.py --configs 3 --cfg3-cells 192 > gpurun_out/r2i_cfg3.json 2> gpurun_out/r2i_cfg3.log; echo "cfg3 exit $?"
tail -5 gpurun_out/r2i_cfg3.log | cut -c1-600; cut -c1-1500 gpurun_out/r2i_cfg3.json
