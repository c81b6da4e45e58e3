#!/bin/bash
# Eight-GPU round: parity of the row-partitioned path at world = 8 (middle ranks with two neighbours), then the
# weak-scaling point of the north star: 512^3 cells in eight z-slabs.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
DIST_CHECK_CASES=0,2,4 timeout 400 $TR --master-port 29521 tools/dist_check.py > gpurun_out/dist_check_n$N.log 2>&1; echo "dist_check exit $?"
grep "DIST_CHECK\|maxrel" gpurun_out/dist_check_n$N.log | cut -c1-300
MGB200_P2P_TRACE=1 timeout 600 $TR --master-port 29522 bench.py --gpus $N > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.log; echo "bench n$N exit $?"
cut -c1-900 gpurun_out/bench_n$N.json
grep -h "slab setup\|upload\|relres" gpurun_out/bench_n$N.log | head -6 | cut -c1-200
grep -h "p2p trace" gpurun_out/bench_n$N.log | sort | head -24
free -g | head -2; nproc
