#!/bin/bash
# one run per variant and workload (values are stable to ~0.5 % run to run)
for wl in ${1:-cfg2 cfg1 cfg5c}; do for v in ${2:-base}; do
  BUDDHA_LIB=$PWD/tools/ab/$v.so timeout -s KILL 200 python bench.py --workload $wl --steps 2 --warmup 1 --skip-baselines --no-extras --samples-per-step 4294967296 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl %-8s %.3e samples/s  %.1f ms/step  exec/S %.2f' % ('$v', d['value'], d['ms_per_step'], d['counters']['executed_iters']/(d['steps']*d['run']['samples_per_step_per_gpu'])))"
done; done
