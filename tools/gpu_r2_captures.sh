#!/bin/bash
# ncu --set full of one render_persistent_kernel launch (2^30 samples) for configs 1-4 with the
# current kernel: raw + source pages as CSV (the .ncu-rep files stay on the box).
mkdir -p gpurun_out
for wl in ${WLS:-cfg1 cfg2 cfg3 cfg4}; do
  BUDDHA_TILE_SERIAL=1 ncu --set full --clock-control none --import-source on -k regex:render_persistent -s 1 -c 1 -f -o /tmp/cap_$wl \
    timeout -s KILL 600 python bench.py --workload $wl --steps 1 --warmup 1 --skip-baselines --no-extras --samples-per-step 1073741824 > /tmp/cap_$wl.log 2>&1
  ncu -i /tmp/cap_$wl.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r2g_${wl}_src.csv
  ncu -i /tmp/cap_$wl.ncu-rep --page raw --csv > gpurun_out/r2g_${wl}_raw.csv
  echo "$wl done: $(tail -c 300 /tmp/cap_$wl.log | tr '\n' ' ' | cut -c1-200)"
done
