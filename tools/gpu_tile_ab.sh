#!/bin/bash
# tiled workloads, serial vs co-resident apply, for several library builds: tools/gpu_tile_ab.sh "cfg3" "p3 pre"
WLS=${1:-cfg3}; VARS=${2:-"p3 pre"}; N=${3:-8589934592}
for wl in $WLS; do for v in $VARS; do for serial in 0 1; do
  if [ $serial = 1 ]; then export BUDDHA_TILE_SERIAL=1; else unset BUDDHA_TILE_SERIAL; fi
  BUDDHA_LIB=$PWD/tools/ab/$v.so timeout -s KILL 200 python bench.py --workload $wl --steps 2 --warmup 1 --skip-baselines --no-extras --samples-per-step $N 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl $v serial=$serial: %.3e samples/s  %.1f ms/step' % (d['value'], d['ms_per_step']))"
done; done; done
unset BUDDHA_TILE_SERIAL
for wl in $WLS; do for v in $VARS; do
BUDDHA_TILE_SERIAL=1 BUDDHA_LIB=$PWD/tools/ab/$v.so ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_${wl}_$v.csv \
  timeout -s KILL 300 python bench.py --workload $wl --steps 1 --warmup 1 --skip-baselines --no-extras --samples-per-step 4294967296 > /dev/null 2>&1
python - "$wl" "$v" <<'PY'
import csv, sys, collections
wl, v = sys.argv[1:3]
rows = [l for l in open('gpurun_out/launches_%s_%s.csv' % (wl, v)) if l.startswith('"')]
r = list(csv.reader(rows))
t = collections.defaultdict(list)
for x in r[1:]:
    t[x[4].split('(')[0].split('<')[0]].append(float(x[-1]) / 1e6)
print(wl, v, "serial launch list:", {k: "%d x %.2f ms (sum %.1f)" % (len(a), sum(a) / len(a), sum(a)) for k, a in t.items() if sum(a) > 0.5})
PY
done; done
