#!/bin/bash
# End-of-milestone evidence run: tests, smoke, default bench + reference arm, workload sweep,
# launch list of the default bench command.
mkdir -p gpurun_out
bash tools/gpu_check.sh tests smoke bench benchref
REF=1 bash tools/gpu_workloads.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_default.csv \
  timeout -s KILL 300 python bench.py --steps 2 --warmup 1 --skip-baselines > gpurun_out/launches_default.log 2>&1
