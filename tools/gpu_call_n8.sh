#!/bin/bash
# 8 GPUs: parity at world 8, cfg2 weak-scaled to 512^3, cfg5 (513^3 ComplexF64) strong-scaled over 8 GPUs.
# COST: an 8-GPU call is charged 8 x its box time, and most of that time is host setup of 8 ranks (the round-2 call took
# 12 minutes = 96 GPU-minutes and was cut before the second step finished).  Run the three steps as separate calls.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/n8_dist_check_n8.log 2>&1; echo "dist_check exit $?"
grep "case\|DIST_CHECK\|Error\|error" gpurun_out/n8_dist_check_n8.log | tail -12 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu > gpurun_out/n8_bench_n8.json 2> gpurun_out/n8_bench_n8.log; echo "bench n8 exit $?"
cut -c1-330 gpurun_out/n8_bench_n8.json
grep "per-kernel" gpurun_out/n8_bench_n8.log | tail -1 | cut -c1-1500
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --config 5 --gpus 8 --steps 10 --warmup 3 > gpurun_out/n8_cfg5_512_n8.json 2> gpurun_out/n8_cfg5_512_n8.log; echo "cfg5 n8 exit $?"
cut -c1-400 gpurun_out/n8_cfg5_512_n8.json; tail -3 gpurun_out/n8_cfg5_512_n8.log | cut -c1-300
