#!/bin/bash
# Runs bench.py on every BASELINE.json workload (1 GPU, short) and collects the JSON lines.
mkdir -p gpurun_out
: > gpurun_out/workloads.jsonl
for wl in ${WLS:-cfg1 cfg2 cfg3 cfg3_m20000 cfg4 cfg5a cfg5b cfg5c}; do
  timeout -s KILL 300 python bench.py --workload $wl --steps 2 --warmup 1 --skip-baselines --no-extras \
    --samples-per-step ${1:-4294967296} >> gpurun_out/workloads.jsonl 2>> gpurun_out/workloads.err
done
if [ -n "${REF:-}" ]; then
  # the unmodified reference binary on the same workloads, -t 3 each (time as the program prints it)
  python - <<'PY' > gpurun_out/workloads_ref.jsonl
import json, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import bench
for wl in "cfg1 cfg2 cfg3 cfg3_m20000 cfg4 cfg5a cfg5b cfg5c".split():
    r = bench.run_reference_binary(bench.WORKLOADS[wl], 3.0)
    print(json.dumps({"workload": wl, "reference_samples_per_s": r[0] if r else None,
                      "passes": r[1] if r else None, "seconds": r[2] if r else None}), flush=True)
PY
  cat gpurun_out/workloads_ref.jsonl
fi
python - <<'PY'
import json
for l in open('gpurun_out/workloads.jsonl'):
    d=json.loads(l); c=d['counters']; S=d['run']['samples_per_step_per_gpu']*d['steps']
    print("%-12s %.3e samples/s  %.3e pts/s  e2e %.3e  frac %.2f  exec/S %.1f ref/S %.1f P/S %.3f exact %.4f ms/step %.1f" % (
        d['config']['workload'].split(':')[0], d['value'], d['orbit_points_per_s'], d['e2e']['value'], d['roofline']['frac'],
        c['executed_iters']/S, c['escape_iters']/S, c['orbit_points']/S, c['exact_bins']/max(c['orbit_points'],1), d['ms_per_step']))
PY
