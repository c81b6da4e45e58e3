#!/bin/bash
# round 2, call o (2 GPUs): which kernel runs on which rank (debug), then the block (multi-RHS) box kernel: parity + cfg4 timing
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
MGB200_DEBUG_KERNELS=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.log; echo "bench n2 exit $?"
grep "mgb200 dev" gpurun_out/r2o_bench_n2.log | sort | uniq -c | sort -rn | head -30
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "block or blockCG or blockFGMRES or blockBiCG or spmatmul" 2>&1 | tail -4
timeout 900 python -m pytest tests/test_baseline_sizes.py -m gpu -q -x -k "cfg4" 2>&1 | tail -4
timeout 900 python tools/bench_configs.py --configs 4 > gpurun_out/r2o_cfg4.json 2> gpurun_out/r2o_cfg4.log; echo "cfg4 exit $?"
cut -c1-1600 gpurun_out/r2o_cfg4.json
