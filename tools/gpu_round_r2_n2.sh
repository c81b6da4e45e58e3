#!/bin/bash
# Two-GPU check of the line-blocked kernels with the fused put (after tools/gpu_round_r2_microbench.sh passed on one GPU):
# parity against the global CPU oracle, then the weak-scaling bench with and without them.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for R in 2 4; do
  MGB200_LINES=$R MGB200_LINES_MIN_ROWS=0 timeout 300 $TR --master-port 2955$R tools/dist_check.py > gpurun_out/dist_check_n${N}_lines$R.log 2>&1; echo "dist_check lines=$R exit $?"
  grep "DIST_CHECK\|rror" gpurun_out/dist_check_n${N}_lines$R.log | cut -c1-200
done
for R in 0 2 4; do
  MGB200_LINES=$R timeout 400 $TR --master-port 2956$R bench.py --gpus $N > gpurun_out/bench_n${N}_lines$R.json 2> gpurun_out/bench_n${N}_lines$R.log; echo "bench lines=$R exit $?"
  cut -c1-300 gpurun_out/bench_n${N}_lines$R.json
done
