#!/bin/bash
# First GPU call of the next round: the line-blocked forms of the sweep and of the transfer operators against the
# one-row-per-thread kernels (bit-compare + time), ~1 minute on one B200.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for t in microbench_lines microbench_transfer; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o gpurun_out/$t tools/$t.cu || exit 1
  timeout 120 gpurun_out/$t > gpurun_out/$t.log 2>&1; echo "$t exit $?"
  cat gpurun_out/$t.log
done
# then the library form of the line-blocked kernel: bit-identity on the GPU, and the bench with it
MGB200_TEST_LINES=1 timeout 300 python -m pytest tests/test_patterns.py -m gpu -q -k line_blocked 2>&1 | tail -3
MGB200_TEST_LINES=1 MGB200_LINES_STAGED=0 timeout 300 python -m pytest tests/test_patterns.py -m gpu -q -k line_blocked 2>&1 | tail -3
for ST in 1 0; do for R in 2 4; do
  MGB200_LINES=$R MGB200_LINES_STAGED=$ST timeout 300 python bench.py --no-cpu > gpurun_out/bench_n1_lines${R}_st$ST.json 2> gpurun_out/bench_n1_lines${R}_st$ST.log; echo "bench lines=$R staged=$ST exit $?"
  cut -c1-200 gpurun_out/bench_n1_lines${R}_st$ST.json
  grep "per-kernel" gpurun_out/bench_n1_lines${R}_st$ST.log | cut -c1-900
done; done
# the grid-hinted transfer kernels: bit-identity on the GPU, and the bench with them
MGB200_TEST_GRID_TRANSFERS=1 timeout 300 python -m pytest tests/test_patterns.py -m gpu -q -k grid_hinted 2>&1 | tail -3
for R in 1 2 4; do
  MGB200_GRID_TRANSFERS=$R timeout 300 python bench.py --no-cpu > gpurun_out/bench_n1_gx$R.json 2> gpurun_out/bench_n1_gx$R.log; echo "bench grid_transfers=$R exit $?"
  cut -c1-200 gpurun_out/bench_n1_gx$R.json
  grep "per-kernel" gpurun_out/bench_n1_gx$R.log | cut -c1-900
done
# full ncu capture of the new kernels inside the bench (edit the options to the winners of the runs above)
MGB200_LINES=4 MGB200_GRID_TRANSFERS=2 timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:'pat_lines|gx_kernel' -c 12 -f -o gpurun_out/lines_gx_full python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/bench_ncu_lines.log 2>&1
echo "ncu exit $?"
