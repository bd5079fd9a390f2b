#!/bin/bash
# Round-2 final evidence on one GPU: parity tests, smoke, the default bench line, the launch list of
# the default command, and one ncu --set full capture of the render kernel (config 2, 2^32 samples).
bash tools/gpu_r2_check.sh 5
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_default.csv \
  timeout -s KILL 300 python bench.py --steps 2 --warmup 1 --skip-baselines --no-extras > gpurun_out/launches_default.log 2>&1
bash tools/gpu_ncu_kernel.sh render_persistent cfg2 4294967296
ncu -i gpurun_out/prof_render_persistent_cfg2.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/final_cfg2_src.csv
ncu -i gpurun_out/prof_render_persistent_cfg2.ncu-rep --page raw --csv > gpurun_out/final_cfg2_raw.csv
