#!/bin/bash
# ncu launch list (device time per launch) for one workload: tools/gpu_launchlist.sh cfg3 [samples]
WL=${1:-cfg3}; N=${2:-1073741824}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$WL.csv \
  timeout -s KILL 300 python bench.py --workload $WL --steps 2 --warmup 1 --skip-baselines --no-extras --samples-per-step $N > gpurun_out/launches_$WL.log 2>&1
python - "$WL" <<'PY'
import csv,sys
rows=[l for l in open('gpurun_out/launches_%s.csv'%sys.argv[1]) if l.startswith('"')]
r=list(csv.reader(rows))
for x in r[1:]:
    if 'buddha' in x[4]: print(x[0], x[4].split('(')[0], x[8], float(x[-1])/1e6, "ms")
PY
