#!/bin/bash
# occupancy sensitivity of the render kernel: pad dynamic shared memory so fewer CTAs fit per SM
run() { echo "== $1 $2"; env BUDDHA_PAD_SMEM=$2 timeout -s KILL 200 python bench.py --workload $1 --steps 2 --warmup 1 --skip-baselines --no-extras --samples-per-step 4294967296 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3e samples/s  %.1f ms/step' % (d['value'], d['ms_per_step']))"; }
for wl in cfg2 cfg1; do for pad in 0 1024 8192 20480; do run $wl $pad; done; done
