#!/bin/bash
# Two-GPU round: parity of the row-partitioned path against the global CPU oracle (peer-memory exchange beside the
# interior rows, with the one-pass and the TMA-staged dictionary kernels), then the weak-scaling bench with and
# without the overlap and on the slab shape of the 512^3 point (64 planes of 513^2 nodes per GPU).
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_n2.log 2>&1; echo "dist_check exit $?"
grep "DIST_CHECK\|maxrel" gpurun_out/dist_check_n2.log | cut -c1-300
MGB200_TMA_MIN_ROWS=0 DIST_CHECK_CASES=0,2,4 timeout 300 $TR --master-port 29512 tools/dist_check.py > gpurun_out/dist_check_n2_tma.log 2>&1; echo "dist_check tma exit $?"
grep "DIST_CHECK\|maxrel" gpurun_out/dist_check_n2_tma.log | cut -c1-300
MGB200_P2P_TRACE=1 timeout 400 $TR --master-port 29513 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.log; echo "bench n2 exit $?"
cut -c1-700 gpurun_out/bench_n2.json
MGB200_OVERLAP=0 timeout 400 $TR --master-port 29514 bench.py --gpus 2 > gpurun_out/bench_n2_nooverlap.json 2> gpurun_out/bench_n2_nooverlap.log; echo "bench n2 no-overlap exit $?"
cut -c1-400 gpurun_out/bench_n2_nooverlap.json
timeout 500 $TR --master-port 29515 bench.py --gpus 2 --grid 512,512,128 > gpurun_out/bench_n2_slab512.json 2> gpurun_out/bench_n2_slab512.log; echo "bench n2 slab512 exit $?"
cut -c1-700 gpurun_out/bench_n2_slab512.json
MGB200_OVERLAP=0 timeout 500 $TR --master-port 29516 bench.py --gpus 2 --grid 512,512,128 > gpurun_out/bench_n2_slab512_nooverlap.json 2> gpurun_out/bench_n2_slab512_nooverlap.log; echo "bench n2 slab512 no-overlap exit $?"
cut -c1-400 gpurun_out/bench_n2_slab512_nooverlap.json
grep -h "p2p trace" gpurun_out/bench_n2.log | head -12
