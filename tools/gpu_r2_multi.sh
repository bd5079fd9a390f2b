#!/bin/bash
# Round-2 multi-GPU check (run under gpurun --gpus N): the 2-GPU in-process tests, then the default
# bench line at N GPUs (strong-scaling job, CLI merge check and all workloads included).
N=${1:-2}; STEPS=${2:-5}
mkdir -p gpurun_out
nvidia-smi -L | head -$N
timeout -s KILL 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "in_process" 2>&1 | tail -4
timeout -s KILL 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps $STEPS --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${N}gpu.err | cut -c1-300
python - $N <<'PY'
import json, sys
d = json.loads(open('gpurun_out/bench_%sgpu.json' % sys.argv[1]).read())
print("N=%d headline %.3e samples/s  e2e %.3e (%.2f of device)  merge %.2f ms  frac %.2f  clocks %s" % (
    d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['value'] / d['value'], d['merge_ms'], d['roofline']['frac'], d['clocks']))
s = d.get('strong_cfg3_m20000')
if s: print("strong: %.3e samples/s  %.0f ms  render max %.0f ms  merge %.1f ms  fnv %s  clk samples %s" % (s['samples_per_s'], s['ms'], s['render_ms_max'], s['merge_ms'], s['hist_fnv'], s['clocks']['samples']))
print("cli_merge:", d.get('cli_merge'))
for w in d.get('workloads', []):
    print("%-6s %.3e samples/s  %.3e pts/s  e2e %.2f of device  merge %.1f ms  steps %d  clk samples %s" % (
        w['workload'], w['value'], w['orbit_points_per_s'], w['e2e']['frac_of_device_rate'], w['merge_ms'], w['steps'], w['clocks']['samples'] if w['clocks'] else None))
PY
