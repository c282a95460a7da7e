import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import cudaraster_linux_b200 as crb
from tests import util
from oracle import binding as G
import ctypes
w, h, n, size, seed = 1920, 1080, 60000, 0.05, 64001
r = crb.CudaRaster(0)
v, i = crb.scenes.random_soup(n, seed=seed, stride_floats=8, size=size)
cc, cd = util.draw_cuda(r, crb, v, i, w, h, "gouraud", 3, 0)
g = util.draw_gold(v, i, w, h, "gouraud", 3, 0)
bad = np.argwhere(cd != g["depth"])
print("flags", os.environ.get("CRB_DEBUG_FLAGS"), "mismatches", len(bad))
if len(bad):
    wb = r.getWorkBuffers(n)
    gs = util.gold_setup(v, i, w, h, "gouraud", 3)
    hdr = gs["triHeader"]; sub = gs["triSubtris"]
    L = G.lib()
    y, x = bad[0]
    tx, ty = x // 8, y // 8
    tiles_x = (w + 7) // 8
    t = ty * tiles_x + tx
    q = wb["tileQueue"][wb["tileStart"][t]:wb["tileStart"][t] + wb["tileCount"][t]]
    print("pixel", x, y, "tile", tx, ty, "queue len", len(q))
    # which triangles cover this pixel according to the oracle?
    bit = (x & 7) + 8 * (y & 7)
    cover = []
    for tri in np.nonzero(sub == 1)[0]:
        m = L.gold_cover_tile(hdr[tri].ctypes.data, w, h, int(tx), int(ty))
        if (m >> bit) & 1:
            cover.append(int(tri))
    print("oracle: triangles covering the pixel:", cover, "in queue:", [int(c * 8 + 7) in set(q.tolist()) for c in cover])
    for c in cover:
        hh = hdr[c]
        xs = [np.int16(hh[k] & 0xFFFF) for k in range(3)]; ys = [np.int16(hh[k] >> 16) for k in range(3)]
        print("  tri", c, "verts(subpx, centred)", list(zip([int(a) for a in xs], [int(b) for b in ys])), "cuda hdr equal:", np.array_equal(wb["triHeader"][c], hh))
