#!/bin/bash
# round 2, call n (2 GPUs): 128-byte aligned padded vectors: parity, bench at N = 2
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r2n_dist_check_n2.log 2>&1; echo "dist_check exit $?"
grep "DIST_CHECK\|Error\|error" gpurun_out/r2n_dist_check_n2.log | tail -5 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.log; echo "bench n2 exit $?"
cut -c1-330 gpurun_out/r2n_bench_n2.json
grep "per-kernel" gpurun_out/r2n_bench_n2.log | cut -c1-1500
timeout 600 python -m pytest tests/test_multi_device.py tests/test_patterns.py -m gpu -q -x 2>&1 | tail -4
