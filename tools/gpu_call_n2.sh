#!/bin/bash
# 2 GPUs: this round's kernels on row-partitioned levels: parity (dist_check, multi-device tests), bench at
# N = 2 (cfg2 weak-scaled) and cfg5 at 256^3 cells strong-scaled over 2 GPUs against the oracle
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_multi_device.py tests/test_patterns.py -m gpu -q -x -k "multi or round2" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r2v_dist_check_n2.log 2>&1; echo "dist_check exit $?"
grep "case\|DIST_CHECK\|Error\|error" gpurun_out/r2v_dist_check_n2.log | tail -12 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2v_bench_n2.json 2> gpurun_out/r2v_bench_n2.log; echo "bench n2 exit $?"
cut -c1-330 gpurun_out/r2v_bench_n2.json
grep "per-kernel" gpurun_out/r2v_bench_n2.log | tail -1 | cut -c1-1800
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --config 5 --cfg5-cells 256 --cfg5-levels 6 --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2v_cfg5_256_n2.json 2> gpurun_out/r2v_cfg5_256_n2.log; echo "cfg5 256 n2 exit $?"
cut -c1-400 gpurun_out/r2v_cfg5_256_n2.json; tail -3 gpurun_out/r2v_cfg5_256_n2.log | cut -c1-300
