#!/bin/bash
# A/B builds of libbuddha.so on the same box: tools/gpu_ab.sh "cfg1 cfg2" "A B C" [samples]
WLS=${1:-cfg2}; VARS=${2:-"A B"}; N=${3:-4294967296}
for wl in $WLS; do for rep in 1 2; do for v in $VARS; do
  BUDDHA_LIB=$PWD/tools/ab/$v.so timeout -s KILL 200 python bench.py --workload $wl --steps 2 --warmup 1 --skip-baselines --no-extras --samples-per-step $N 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl $v: %.3e samples/s  %.1f ms/step  e2e %.3e  exec/S %.2f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['counters']['executed_iters']/(d['steps']*d['run']['samples_per_step_per_gpu'])))"
done; done; done
