#!/bin/bash
# ncu --set full on one kernel: tools/gpu_ncu_kernel.sh <kernel regex> <workload> <samples> [skip]
K=$1; WL=${2:-cfg3}; N=${3:-1073741824}; S=${4:-1}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c 1 -f -o gpurun_out/prof_${K}_$WL \
  timeout -s KILL 600 python bench.py --workload $WL --steps 1 --warmup 1 --skip-baselines --no-extras --samples-per-step $N > gpurun_out/ncu_$K.log 2>&1
tail -2 gpurun_out/ncu_$K.log
