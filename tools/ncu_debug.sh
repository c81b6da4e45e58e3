#!/bin/bash
# Diagnose the LaunchFailed that ncu reports on the level-2 (27-point) TMA-staged dictionary kernel.
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
export PATH=/usr/local/cuda/bin:$PATH
echo "== memcheck, TMA forced on every level, 64^3"
MGB200_TMA_MIN_ROWS=0 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python bench.py --cells 64 --levels 4 --steps 1 --warmup 1 --no-cpu > gpurun_out/memcheck.log 2>&1
echo "exit $?"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/memcheck.log; grep -m 12 "Invalid\|ERROR SUMMARY\|at \|by thread\|Address" gpurun_out/memcheck.log
echo "== racecheck"
MGB200_TMA_MIN_ROWS=0 timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python bench.py --cells 48 --levels 3 --steps 1 --warmup 1 --no-cpu > gpurun_out/racecheck.log 2>&1
echo "exit $?"; grep -m 12 "RACECHECK\|hazard\|ERROR" gpurun_out/racecheck.log
for variant in graphs0 cachenone default; do
  echo "== ncu $variant (128^3, level 2 = 65^3 rows uses the TMA kernel)"
  extra=""; env_g=1
  [ $variant = graphs0 ] && env_g=0
  [ $variant = cachenone ] && extra="--cache-control none"
  MGB200_GRAPHS=$env_g timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none $extra --profile-from-start off -c 80 --csv \
      --log-file gpurun_out/dbg_launches_$variant.csv python bench.py --cells 128 --levels 5 --steps 2 --warmup 3 --no-cpu > gpurun_out/dbg_ncu_$variant.log 2>&1
  echo "exit $?  rows $(grep -c gpu__time_duration gpurun_out/dbg_launches_$variant.csv)"
  grep -m 3 "ERROR" gpurun_out/dbg_launches_$variant.csv gpurun_out/dbg_ncu_$variant.log
done
