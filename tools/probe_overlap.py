#!/usr/bin/env python
"""Developer probe: C2 frames rendered by ONE context on one stream vs TWO independent contexts on two streams (two views in
flight on one GPU).  Setup is issue-bound and the fine raster latency-bound, so concurrent frames can fill each other's gaps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cudaraster_linux_b200 as crb  # noqa: E402

w, h, steps = 1920, 1080, 40
verts, idx = crb.scenes.grid_gouraud(1000, 500)
dev = torch.device("cuda", 0)
n_ctx = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctxs = []
for c in range(n_ctx):
    r = crb.CudaRaster(0)
    color = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8)
    depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32)
    copies = [(torch.from_numpy(verts).to(dev), torch.from_numpy(idx).to(dev)) for _ in range(2)]
    s = torch.cuda.Stream(device=dev)
    r.setSurfaces(color, depth)
    r.setPixelPipe(None, crb.pipe_name("gouraud", 0, 3))
    ctxs.append((r, color, depth, copies, s))


def frame(c, k, asynchronous=True):
    r, color, depth, copies, s = ctxs[c]
    vb, ib = copies[k % 2]
    r.setVertexBuffer(vb, 0)
    r.setIndexBuffer(ib, 0, idx.shape[0])
    r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
    r.drawTriangles(stream=s.cuda_stream, asynchronous=asynchronous)


for c in range(n_ctx):
    for k in range(4):
        frame(c, k, asynchronous=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
main = torch.cuda.current_stream()
e0.record(main)
for c in range(n_ctx):
    ctxs[c][4].wait_stream(main)
for k in range(steps):
    for c in range(n_ctx):
        frame(c, k)
for c in range(n_ctx):
    main.wait_stream(ctxs[c][4])
e1.record(main)
for c in range(n_ctx):
    ctxs[c][0].finish(stream=ctxs[c][4].cuda_stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print("contexts", n_ctx, "frames", steps * n_ctx, "ms/frame %.4f" % (ms / (steps * n_ctx)), "Mtris/s %.0f" % (idx.shape[0] * steps * n_ctx / ms / 1e3))
