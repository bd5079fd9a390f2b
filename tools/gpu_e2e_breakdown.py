import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import cudabrot_b200 as B
r = B.Renderer(4000, 4000, 20000, 10000)
cells = 16000000
hh = torch.zeros(cells, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
hi = torch.zeros(cells, dtype=torch.int16).pin_memory().numpy().view(np.uint16)
for step in range(3):
    t0 = time.perf_counter(); r.load_histogram(hh); t1 = time.perf_counter()
    r.render_samples(step << 40, 1 << 35); t2 = time.perf_counter()
    r.read_histogram(hh.reshape(4000, 4000)); t3 = time.perf_counter()
    _, mx, _ = r.tonemap(1.0, True, out=hi.reshape(4000, 4000)); t4 = time.perf_counter()
    print("load %.1f ms  render %.1f ms (dev %.1f)  read %.1f ms  tonemap %.1f ms (kernels %.3f)  max %d" % (
        1e3*(t1-t0), 1e3*(t2-t1), r.last_render_ms(), 1e3*(t3-t2), 1e3*(t4-t3), r.last_tonemap_ms(), mx), flush=True)
