#!/bin/bash
# compute-sanitizer over small renders: memcheck (global/shared OOB), racecheck (shared-memory
# hazards in the warp-synchronous stacks), synccheck.  Plain, tiled, fused and ship variants.
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys, numpy as np
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
run(width=64, height=48, max_iterations=400, min_iterations=20, n=1 << 16)   # privatised copies
# round 2: the cycle certificate (queues in global memory, per-lane refill, re-injection into deep)
os.environ["BUDDHA_CERT_QUEUE"] = "160"
run(width=200, height=150, max_iterations=10000, min_iterations=20, n=1 << 18)
del os.environ["BUDDHA_CERT_QUEUE"]
# round 2: a pipeline of tiled launches with orbit carry-over (1 MB list pool -> many launches)
import os
os.environ["BUDDHA_TILE_POOL_MB"] = "1"
run(width=300, height=200, max_iterations=2000, min_iterations=20, flags=B.F_FORCE_TILED, n=1 << 19)
del os.environ["BUDDHA_TILE_POOL_MB"]
# round 2: overlapped transfers, digest, colour combine
with B.Renderer(160, 120, channels=[(100, 20), (1000, 20), (3000, 20)]) as r:
    saved = np.ones((3, 120, 160), dtype=np.uint32)
    r.add_histogram_async(saved); r.render_samples_async(0, 1 << 16); r.snapshot()
    r.render_samples_async(1 << 16, 1 << 16)
    s = r.read_snapshot(); img, mx, sc = r.tonemap_snapshot(2.2, True, channel=1); r.sync()
    rgb, m3 = r.combine_rgb((0, 1, 2), gamma=2.2, mode="hsl", hue_adjust=0.3)
    print("async", int(s.sum()), mx, hex(r.digest(2)), m3)
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"
  timeout -s KILL 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/san.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Error|error|hazard|^[0-9]|^async" | head -20
done
