mkdir -p gpurun_out
tools/gpu_ncu_kernel.sh render_persistent cfg2 1073741824 1
tools/gpu_ncu_kernel.sh render_persistent cfg1 1073741824 1
tools/gpu_ncu_kernel.sh render_persistent cfg4 1073741824 1
BUDDHA_TILE_SERIAL=1 tools/gpu_ncu_kernel.sh apply_tile cfg3 1073741824 30
BUDDHA_TILE_SERIAL=1 tools/gpu_ncu_kernel.sh render_persistent cfg3 1073741824 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_default.csv \
  timeout -s KILL 300 python bench.py --steps 2 --warmup 1 --skip-baselines --no-extras > gpurun_out/launches_default.log 2>&1
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read())
print("headline %.3e samples/s  e2e %.3e (%.2f of device)  frac %.2f  launches %d  clocks %s" % (
    d['value'], d['e2e']['value'], d['e2e']['value'] / d['value'], d['roofline']['frac'], d['gpu_launches'], d['clocks']))
s = d.get('strong_cfg3_m20000')
if s: print("strong: %.3e samples/s  %.0f ms  merge %.1f ms  fnv %s" % (s['samples_per_s'], s['ms'], s['merge_ms'], s['hist_fnv']))
for w in d.get('workloads', []):
    print("%-6s %.3e samples/s  %.3e pts/s  e2e %.2f of device  fp64 frac %.2f  red frac %.2f  steps %d  clk samples %s  vs ref %s" % (
        w['workload'], w['value'], w['orbit_points_per_s'], w['e2e']['frac_of_device_rate'], w['roofline']['frac'],
        w['roofline_red']['frac'], w['steps'], w['clocks']['samples'] if w['clocks'] else None, w.get('vs_reference_cuda')))
PY
