#!/bin/bash
run() { echo "== $*"; env "$@" timeout -s KILL 200 python bench.py --workload cfg3 --steps 2 --warmup 1 --skip-baselines --no-extras --samples-per-step 4294967296 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e samples/s  %.1f ms/step  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"; }
run A=1
run BUDDHA_TILE_SERIAL=1
run BUDDHA_TILE_LAUNCH_LOG2=28
run BUDDHA_TILE_LAUNCH_LOG2=29
run BUDDHA_TILE_MIN_MB=100000
run BUDDHA_TILE_SHIFT=23
run BUDDHA_TILE_SHIFT=25
