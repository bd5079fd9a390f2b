"""Roofline probes on one B200: FP64-pipe issue rate and red.global.add.u32 rate vs footprint."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabrot_b200 as B
out = {}
with B.Renderer(64, 64, 10, 0) as r:
    out["fp64_lane_instr_per_s"] = r.probe_fp64_peak()
    out["red_per_s"] = {}
    for mb in (4, 64, 128, 400, 1600, 6400):
        out["red_per_s"]["%d MB" % mb] = r.probe_red_peak(mb * 1000 * 1000)
print(json.dumps(out))
