#!/bin/bash
run() { echo "== $*"; env "$@" bash -c 'timeout -s KILL 200 python bench.py --workload $WL --steps 2 --warmup 1 --skip-baselines --no-extras --samples-per-step 4294967296 2>/dev/null' | python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['counters']; print('%.3e samples/s  %.3e pts/s  %.1f ms/step exec/S %.2f' % (d['value'], d['orbit_points_per_s'], d['ms_per_step'], c['executed_iters']/(d['steps']*d['run']['samples_per_step_per_gpu'])))"; }
run WL=cfg2
run WL=cfg5c
for wl in cfg5a cfg5b cfg5c; do run WL=$wl BUDDHA_TILE_MIN_MB=256; done
run WL=cfg5a BUDDHA_TILE_MIN_MB=256 BUDDHA_TILE_SHIFT=23
