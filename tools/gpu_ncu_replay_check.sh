#!/bin/bash
# Does bench.py's "histogram sum == increments" self-check hold under ncu's multi-pass kernel replay?
# (base = the build before the cycle certificate, new with the certificate off / on)
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" ncu --set full --clock-control none -k regex:render_persistent -s 1 -c 1 -f -o /tmp/rep_$name \
    timeout -s KILL 600 python bench.py --workload cfg2 --steps 1 --warmup 1 --skip-baselines --no-extras --samples-per-step 4294967296 > /tmp/rep_$name.log 2>&1
  echo "$name: $(grep -c 'histogram sum' /tmp/rep_$name.log) mismatch lines; $(grep -o 'histogram sum [0-9]* != increments [0-9]*' /tmp/rep_$name.log | head -1)"
}
run base BUDDHA_LIB=$PWD/tools/ab/base.so
run new_off BUDDHA_CERT_QUEUE=0
run new_on X=1
# without the profiler, the certificate on, small queues: three runs
for i in 1 2 3; do BUDDHA_CERT_QUEUE=160 timeout -s KILL 300 python bench.py --workload cfg2 --steps 2 --warmup 1 --skip-baselines --no-extras --samples-per-step 4294967296 2>&1 | grep -o 'histogram sum.*\|"value": [0-9.e+]*' | head -1; done
