#!/bin/bash
# Round-2 final evidence on one GPU: parity tests, smoke, the default bench line, the reference arm,
# and the launch list of the default command.  (ncu --set full captures: tools/gpu_r2_captures.sh)
bash tools/gpu_r2_check.sh 5
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_default.csv \
  timeout -s KILL 300 python bench.py --steps 2 --warmup 1 --skip-baselines --no-extras > gpurun_out/launches_default.log 2>&1
