#!/bin/bash
# 1 GPU: the state at the end of the round - full GPU test suite, smoke, bench line, launch list, ncu of the
# top kernels, the other BASELINE configs
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r2x_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2x_bench_cfg2_n1.json 2> gpurun_out/r2x_bench_cfg2_n1.log; echo "bench exit $?"
cut -c1-400 gpurun_out/r2x_bench_cfg2_n1.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2x_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2x_bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r2x_launches_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1; echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'box_kernel|gxp_quad_kernel|gxr_kernel' --launch-count 24 -f -o /tmp/r2x python tools/ncu_cycle.py > gpurun_out/r2x_ncu.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/r2x.ncu-rep --page raw --csv > gpurun_out/r2x_ncu_raw.csv 2>/dev/null
timeout 1200 python tools/bench_configs.py --configs 1,3,4 --cfg3-cells 192 --out gpurun_out/r2x_configs_1_3_4.json > gpurun_out/r2x_configs.log 2>&1; echo "configs exit $?"; cut -c1-700 gpurun_out/r2x_configs.log
ls -la gpurun_out/r2x*
