#!/bin/bash
# compute-sanitizer over small renders: memcheck (global/shared OOB), racecheck (shared-memory
# hazards in the warp-synchronous stacks), synccheck.  Plain, tiled, fused and ship variants.
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import cudabrot_b200 as B
def run(**kw):
    n = kw.pop('n', 1 << 17)
    with B.Renderer(**kw) as r:
        r.render_samples(3, n)
        h = r.read_histogram(); c = r.counters()
        r.tonemap(2.2, True, channel=0)
        print(kw.get('flags', 0), kw.get('channels'), int(h.sum()), c['increments'], c['accepted'])
run(width=200, height=150, max_iterations=3000, min_iterations=20)
run(width=200, height=150, max_iterations=300, min_iterations=5, flags=B.F_FORCE_TILED)
run(width=200, height=150, channels=[(100, 20), (1000, 20), (3000, 50)])
run(width=200, height=150, channels=[(100, 20), (1000, 20)], flags=B.F_FORCE_TILED | B.F_BURNING_SHIP)
run(width=64, height=64, max_iterations=5, min_iterations=0, n=5000)
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout -s KILL 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|^[0-9]" | head -20
done
