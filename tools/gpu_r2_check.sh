#!/bin/bash
# Round-2 check on the GPU box: parity tests, smoke, the full default bench line (N = 1).
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -12 gpurun_out/pytest_gpu.txt
timeout -s KILL 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout -s KILL 900 python bench.py --steps ${1:-5} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err | cut -c1-400
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read())
print("headline %.3e samples/s  e2e %.3e (%.2f of device)  frac %.2f  launches %d  clocks %s" % (
    d['value'], d['e2e']['value'], d['e2e']['value'] / d['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks']))
s = d.get('strong_cfg3_m20000')
if s: print("strong: %.3e samples/s  %.0f ms  merge %.1f ms  fnv %s  clocks %s" % (s['samples_per_s'], s['ms'], s['merge_ms'], s['hist_fnv'], s['clocks']))
print("cli_merge:", d.get('cli_merge'))
for w in d.get('workloads', []):
    print("%-6s %.3e samples/s  %.3e pts/s  e2e %.2f of device  fp64 frac %.2f  red frac %.2f  steps %d  clk samples %s  vs ref %s" % (
        w['workload'], w['value'], w['orbit_points_per_s'], w['e2e']['frac_of_device_rate'], w['roofline']['frac'],
        w['roofline_red']['frac'], w['steps'], w['clocks']['samples'] if w['clocks'] else None, w.get('vs_reference_cuda')))
print("cpu_baseline", d.get('cpu_baseline', {}).get('value'), "reference_cuda", d.get('reference_cuda', {}).get('value'))
PY
