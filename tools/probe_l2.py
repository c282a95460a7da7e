#!/usr/bin/env python
"""Developer tool: C2-style frames of different size (same triangle size in pixels) -- does the per-pixel cost of the stages drop
when the frame's working set fits the L2?  Prints stage times (asynchronous frames, stage events) per size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cudaraster_linux_b200 as crb
r = crb.CudaRaster(0)
for scale in (4, 2, 1):
    nx, ny, w, h = 1000 // scale, 500 // scale, 1920 // scale, 1080 // scale
    v, i = crb.scenes.grid_gouraud(nx, ny)
    color = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8); depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32)
    vb, ib = torch.from_numpy(v).cuda(), torch.from_numpy(i).cuda()
    r.setSurfaces(color, depth); r.setPixelPipe(None, crb.pipe_name("gouraud", 0, 3)); r.setVertexBuffer(vb, 0); r.setIndexBuffer(ib, 0, i.shape[0])
    for _ in range(3):
        r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0); r.drawTriangles()
    r.setStageTiming(True)
    for _ in range(40):
        r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0); r.drawTriangles(asynchronous=True)
    r.finish(); torch.cuda.synchronize()
    st = r.getStageTiming(); r.setStageTiming(False)
    mpx = w * h / 1e6
    print("scale 1/%d: %d tris %dx%d  setup %.1f us (%.1f us/Mtri)  fine %.1f us (%.1f us/Mpx)  alloc %.1f scatter %.1f" % (
        scale, i.shape[0], w, h, st["triangleSetup"] * 1e3, st["triangleSetup"] * 1e3 / (i.shape[0] / 1e6), st["fineRaster"] * 1e3, st["fineRaster"] * 1e3 / mpx, st["binRaster"] * 1e3, st["coarseRaster"] * 1e3), flush=True)
