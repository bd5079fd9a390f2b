#!/bin/bash
# ms/step of one workload at several step sizes (launch overhead / tail / drain share)
WL=${1:-cfg2}
for lg in 28 30 32 34; do
  n=$((1<<lg))
  timeout -s KILL 300 python bench.py --workload $WL --steps 3 --warmup 2 --skip-baselines --no-extras --samples-per-step $n 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('2^$lg: %.3e samples/s  %.2f ms/step  e2e %.3e' % (d['value'], d['ms_per_step'], d['e2e']['value']))"
done
