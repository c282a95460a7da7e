#!/usr/bin/env python
"""Developer tool: general path (mode 0) vs direct tile path (mode 2) on scenes with LARGE triangles -- the case the automatic
binning mode used to keep away from the direct path.  Prints ms per frame (30 asynchronous frames, CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cudaraster_linux_b200 as crb

def scene(name):
    S = crb.scenes
    if name == "soup200k_mixed":
        return S.random_soup(200000, seed=5, stride_floats=8, size=0.2, clip_fraction=0.05, behind_fraction=0.01)
    if name == "huge2k":
        return S.random_soup(2000, seed=6, stride_floats=8, size=1.5, clip_fraction=0.3, behind_fraction=0.0)
    if name == "mid50k_100px":
        return S.random_soup(50000, seed=7, stride_floats=8, size=0.25, clip_fraction=0.0, behind_fraction=0.0)
    if name == "c2+4fullscreen":
        v, i = S.grid_gouraud(1000, 500)
        big = np.array([[-3, -3, 0.9, 1, 1, 0, 0, 1], [3, -3, 0.9, 1, 0, 1, 0, 1], [0, 3, 0.9, 1, 0, 0, 1, 1]], np.float32)
        n = v.shape[0]
        v = np.concatenate([v] + [big] * 4)
        i = np.concatenate([np.array([[n + 3 * k, n + 3 * k + 1, n + 3 * k + 2] for k in range(4)], np.int32), i])
        return v, i
    if name == "c2":
        return S.grid_gouraud(1000, 500)
    raise KeyError(name)

r = crb.CudaRaster(0)
w, h = 1920, 1080
color = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8)
depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32)
for name in sys.argv[1:] or ["c2", "c2+4fullscreen", "soup200k_mixed", "mid50k_100px", "huge2k"]:
    v, i = scene(name)
    vb, ib = torch.from_numpy(v).cuda(), torch.from_numpy(i).cuda()
    res = {}
    for mode in (0, 2, 1):
        r.setBinningMode(mode)
        r.setSurfaces(color, depth); r.setPixelPipe(None, crb.pipe_name("gouraud", 0, 3)); r.setVertexBuffer(vb, 0); r.setIndexBuffer(ib, 0, i.shape[0])
        for _ in range(3):
            r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0); r.drawTriangles()
        ref = depth.numpy().copy() if mode == 0 else ref
        assert np.array_equal(depth.numpy(), ref), "depth differs between paths"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(30):
            r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0); r.drawTriangles(asynchronous=True)
        e1.record(); r.finish(); torch.cuda.synchronize()
        c = r.getCounters()
        r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0); r.drawTriangles()
        st = r.getStats()
        res[mode] = (round(e0.elapsed_time(e1) / 30, 4), r.lastFrameDirect(), c["numTileEntries"], c["numLargeTris"], [round(st[k] * 1e3, 3) for k in ("setupTime", "binTime", "coarseTime", "fineTime")])
    print(name, i.shape[0], "tris | general", res[0], "| direct", res[2], "| auto", res[1], flush=True)
