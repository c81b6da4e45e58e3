#!/bin/bash
# Two-GPU round after the window merge / alignment changes: parity (default and TMA tiles forced), bench.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29541 tools/dist_check.py > gpurun_out/dist_check_n2_r01i.log 2>&1; echo "dist_check exit $?"
grep "DIST_CHECK\|rror" gpurun_out/dist_check_n2_r01i.log | cut -c1-200
MGB200_TMA_MIN_ROWS=0 DIST_CHECK_CASES=0,1,2,4 timeout 300 $TR --master-port 29542 tools/dist_check.py > gpurun_out/dist_check_n2_r01i_tma.log 2>&1; echo "dist_check tma exit $?"
grep "DIST_CHECK\|rror" gpurun_out/dist_check_n2_r01i_tma.log | cut -c1-200
timeout 400 $TR --master-port 29543 bench.py --gpus $N > gpurun_out/bench_n2_r01i.json 2> gpurun_out/bench_n2_r01i.log; echo "bench exit $?"
cut -c1-330 gpurun_out/bench_n2_r01i.json
