#!/bin/bash
run() { echo "== $*"; env "$@" bash -c 'timeout -s KILL 200 python bench.py --workload $WL --steps 2 --warmup 1 --skip-baselines --no-extras --samples-per-step 4294967296 2>/dev/null' | python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['counters']; print('%.3e samples/s  %.3e pts/s  %.1f ms/step e2e %.3e' % (d['value'], d['orbit_points_per_s'], d['ms_per_step'], d['e2e']['value']))"; }
run WL=cfg3
run WL=cfg3_m20000
run WL=cfg5c BUDDHA_TILE_MIN_MB=256
