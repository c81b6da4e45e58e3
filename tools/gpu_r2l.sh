#!/bin/bash
# round 2, call l (2 GPUs): single-process multi-GPU entry, process-per-GPU parity (dist_check), bench at N = 2
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi -L
timeout 600 python -m pytest tests/test_multi_device.py -m gpu -q -x 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r2l_dist_check_n2.log 2>&1; echo "dist_check exit $?"
grep -v "^\[W\|^W1\|NCCL\|Warning" gpurun_out/r2l_dist_check_n2.log | tail -12 | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2l_bench_n2.json 2> gpurun_out/r2l_bench_n2.log; echo "bench n2 exit $?"
cut -c1-900 gpurun_out/r2l_bench_n2.json; grep "per-kernel" gpurun_out/r2l_bench_n2.log | cut -c1-1500
