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
