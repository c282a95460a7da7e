"""Developer probe: one scene, CUDA vs oracle, prints where they differ."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import cudaraster_linux_b200 as crb
from tests import util

def run(w, h, n, s_log2, size, seed, fresh=True, r=None):
    r = r or crb.CudaRaster(0)
    v, i = crb.scenes.random_soup(max(n, 1), seed=seed, stride_floats=8, size=size)
    cc, cd = util.draw_cuda(r, crb, v, i, w, h, "gouraud", 3, s_log2)
    g = util.draw_gold(v, i, w, h, "gouraud", 3, s_log2)
    bad = np.argwhere(cd != g["depth"])
    print("case", (w, h, n, s_log2, size, seed), "depth mismatches:", len(bad), "colour maxdiff:", util.color_max_diff(cc, g["color"]), "counters", r.getCounters())
    if len(bad):
        print("  bbox y[%d..%d] x[%d..%d]" % (bad[:, 0].min(), bad[:, 0].max(), bad[:, 1].min(), bad[:, 1].max()), "first", bad[:6].tolist())
        ty, tx = bad[:, 0] // 8, bad[:, 1] // 8
        tiles = set(zip(ty.tolist(), tx.tolist()))
        print("  tiles affected:", len(tiles), sorted(tiles)[:10])
        for y, x in bad[:4]:
            print("   (%d,%d) cuda %08x gold %08x" % (x, y, cd[y, x], g["depth"][y, x]))
    return r

r = run(1920, 1080, 60000, 0, 0.05, 4000 + 60000)
run(1920, 1080, 60000, 0, 0.05, 4000 + 60000, r=r)
run(1920, 1080, 20000, 0, 0.05, 7)
run(1920, 1080, 60000, 0, 0.02, 8)
run(1920, 1080, 60000, 0, 0.1, 9)

print("==== sequence of the failing test")
cases = [(640, 360, 20000, 0, 0.3), (320, 200, 3000, 0, 0.8), (640, 360, 20000, 2, 0.3), (1920, 1080, 60000, 0, 0.05), (640, 360, 0, 0, 0.3),
         (640, 360, 25000, 0, 1.5), (320, 200, 3000, 0, 0.8)]
r = crb.CudaRaster(0)
for rep in range(2):
    for w, h, n, s_log2, size in cases:
        if n == 0:
            continue
        run(w, h, n, s_log2, size, 4000 + n + rep, r=r)
print("==== MSAA then 1080p only")
r = crb.CudaRaster(0)
run(640, 360, 20000, 2, 0.3, 24000, r=r)
run(1920, 1080, 60000, 0, 0.05, 64000, r=r)
print("==== small then 1080p only")
r = crb.CudaRaster(0)
run(640, 360, 20000, 0, 0.3, 24000, r=r)
run(1920, 1080, 60000, 0, 0.05, 64000, r=r)
