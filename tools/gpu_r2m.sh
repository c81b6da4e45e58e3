#!/bin/bash
# round 2, call m (2 GPUs): overlapped exchange + slab transfers: parity (dist_check, multi-device tests), bench at N = 2 with and without
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_multi_device.py -m gpu -q -x 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r2m_dist_check_n2.log 2>&1; echo "dist_check exit $?"
grep "case\|DIST_CHECK\|Error\|error" gpurun_out/r2m_dist_check_n2.log | tail -12 | cut -c1-300
for ov in 1 0; do
MGB200_OVERLAP_BOX=$ov timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2m_bench_n2_ov$ov.json 2> gpurun_out/r2m_bench_n2_ov$ov.log; echo "bench n2 overlap=$ov exit $?"
cut -c1-330 gpurun_out/r2m_bench_n2_ov$ov.json
done
MGB200_GRID_TRANSFERS=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu > gpurun_out/r2m_bench_n2_gx0.json 2> gpurun_out/r2m_bench_n2_gx0.log; echo "bench n2 gx=0 exit $?"
cut -c1-330 gpurun_out/r2m_bench_n2_gx0.json
grep "per-kernel" gpurun_out/r2m_bench_n2_ov1.log | tail -1 | cut -c1-1800
