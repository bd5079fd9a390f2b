#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, smoke, a short bench and the ncu passes.
# Usage: tools/gpu_check.sh [tests] [bench] [ncu]
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
for what in "$@"; do
case $what in
tests)
  timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -25 gpurun_out/pytest_gpu.txt ;;
smoke)
  timeout -s KILL 120 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.txt ;;
bench)
  timeout -s KILL 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
  tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json ;;
benchref)
  timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  cat gpurun_out/bench_ref.json ;;
ncu)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv \
      --log-file gpurun_out/launches.csv timeout -s KILL 300 python bench.py --steps 2 --warmup 1 --skip-baselines --no-extras \
      --samples-per-step 1073741824 > gpurun_out/ncu_bench.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:render_persistent -s 1 -c 1 \
      -f -o gpurun_out/prof_render timeout -s KILL 600 python bench.py --steps 1 --warmup 1 --skip-baselines --no-extras \
      --samples-per-step 1073741824 > gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log ;;
esac
done
