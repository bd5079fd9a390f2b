import sys, os, time
sys.path.insert(0, '.')
import cudabrot_b200 as B
def run(side, m, c, n=1<<31):
    with B.Renderer(side, side, m, c) as r:
        r.render_samples(1 << 50, 1 << 28)   # warm-up (+ calibration when tiled)
        r.render_samples(0, n)
        ms = r.last_render_ms()
    return n / ms / 1e-3
for side in (12000, 14000, 16000):
    for (m, c) in ((2000, 20), (20000, 20), (100, 20)):
        out = []
        for mode in ("100000", "1"):
            os.environ["BUDDHA_TILE_MIN_MB"] = mode
            out.append(run(side, m, c))
        print("%5d^2 (%4d MB) m=%5d: direct %.3e  tiled %.3e  -> %s" % (side, side*side*4>>20, m, out[0], out[1], "tiled" if out[1] > out[0] else "direct"), flush=True)
