#!/bin/bash
# round 2, call j: bench.py (N = 1) with the parity record, cfg5 smoke at 128^3 with the oracle, cfg5 at 512^3 on one GPU
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 900 python bench.py > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.log; echo "bench exit $?"
cut -c1-2500 gpurun_out/r2j_bench_n1.json; tail -3 gpurun_out/r2j_bench_n1.log | cut -c1-400
timeout 600 python tools/bench_cfg5.py --cells 128 --levels 5 --steps 5 > gpurun_out/r2j_cfg5_128.json 2> gpurun_out/r2j_cfg5_128.log; echo "cfg5-128 exit $?"
cut -c1-1200 gpurun_out/r2j_cfg5_128.json; tail -4 gpurun_out/r2j_cfg5_128.log | cut -c1-300
free -g | head -2
timeout 1500 python tools/bench_cfg5.py --cells 512 --levels 7 --steps 5 --norms-out gpurun_out/r02_cfg5_n1_norms.json > gpurun_out/r2j_cfg5_512_n1.json 2> gpurun_out/r2j_cfg5_512_n1.log; echo "cfg5-512 exit $?"
cut -c1-3000 gpurun_out/r2j_cfg5_512_n1.json; tail -6 gpurun_out/r2j_cfg5_512_n1.log | cut -c1-300
