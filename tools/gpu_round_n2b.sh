#!/bin/bash
# Two-GPU round for the fused put: parity with the put from the one-pass and the TMA-staged kernels (V, W, K cycles,
# Krylov drivers), then the bench with and without it.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29531 tools/dist_check.py > gpurun_out/dist_check_n${N}_fused.log 2>&1; echo "dist_check exit $?"
grep "DIST_CHECK\|maxrel\|rror" gpurun_out/dist_check_n${N}_fused.log | cut -c1-260
MGB200_TMA_MIN_ROWS=0 DIST_CHECK_CASES=0,1,2,4 timeout 300 $TR --master-port 29532 tools/dist_check.py > gpurun_out/dist_check_n${N}_fused_tma.log 2>&1; echo "dist_check tma exit $?"
grep "DIST_CHECK\|maxrel\|rror" gpurun_out/dist_check_n${N}_fused_tma.log | cut -c1-260
MGB200_P2P_TRACE=1 timeout 400 $TR --master-port 29533 bench.py --gpus $N > gpurun_out/bench_n${N}_fused.json 2> gpurun_out/bench_n${N}_fused.log; echo "bench fused exit $?"
cut -c1-330 gpurun_out/bench_n${N}_fused.json
if [ "$N" = "2" ]; then
MGB200_FUSED_PUT=0 timeout 400 $TR --master-port 29534 bench.py --gpus $N > gpurun_out/bench_n${N}_nofused.json 2> gpurun_out/bench_n${N}_nofused.log; echo "bench no-fused exit $?"
cut -c1-330 gpurun_out/bench_n${N}_nofused.json
timeout 500 $TR --master-port 29535 bench.py --gpus 2 --grid 512,512,128 > gpurun_out/bench_n2_slab512_fused.json 2> gpurun_out/bench_n2_slab512_fused.log; echo "bench slab512 fused exit $?"
cut -c1-330 gpurun_out/bench_n2_slab512_fused.json
fi
grep -h "p2p trace" gpurun_out/bench_n${N}_fused.log | sort | head -12
