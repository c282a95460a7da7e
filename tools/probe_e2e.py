"""Developer probe: where does the pipelined host path spend its time? (not product, not a test)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import cudaraster_linux_b200 as crb
import bench

desc, verts, idx, w, h, shader, s_log2, flags, k_var = bench.make_scene("c2")
dev = torch.device("cuda", 0)
hv, hi = torch.from_numpy(verts).pin_memory(), torch.from_numpy(idx).pin_memory()
dv, di = torch.empty_like(hv, device=dev), torch.empty_like(hi, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s1 = torch.cuda.Stream()
for name, fn in (("H2D verts+idx (torch copy_, side stream)", lambda: (dv.copy_(hv, non_blocking=True), di.copy_(hi, non_blocking=True))),):
    with torch.cuda.stream(s1):
        fn(); torch.cuda.synchronize()
        e0.record(s1)
        for _ in range(10): fn()
        e1.record(s1)
    torch.cuda.synchronize()
    print(name, e0.elapsed_time(e1) / 10, "ms", (hv.numel() * 4 + hi.numel() * 4) / (e0.elapsed_time(e1) / 10 * 1e-3) / 1e9, "GB/s")

r = crb.CudaRaster(0)
color = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8, 1, device=dev)
depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32, 1, device=dev)
r.setSurfaces(color, depth)
r.setPixelPipe(None, crb.pipe_name(shader, s_log2, flags, "BlendReplace"))
hc = torch.zeros_like(color.tensor, device="cpu").pin_memory()
for _ in range(2):
    r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0); r.drawTrianglesHost(hv, hi, idx.shape[0], hc)
torch.cuda.synchronize()
stream = torch.cuda.current_stream(dev)
for n in (1, 2, 4, 8, 30):
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(n):
        r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0); r.drawTrianglesHostAsync(hv, hi, idx.shape[0], hc)
    t1 = time.perf_counter()
    e1.record(stream)
    r.finish(); torch.cuda.synchronize()
    t2 = time.perf_counter()
    print("pipelined n=%d: %.3f ms/frame (events), host enqueue %.3f ms total, wall %.3f ms/frame" % (n, e0.elapsed_time(e1) / n, (t1 - t0) * 1e3, (t2 - t0) * 1e3 / n))
s2 = torch.cuda.Stream()
for n in (8,):
    e0.record(s2)
    for _ in range(n):
        r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0); r.drawTrianglesHostAsync(hv, hi, idx.shape[0], hc, stream=s2.cuda_stream)
    e1.record(s2)
    r.finish(stream=s2.cuda_stream); torch.cuda.synchronize()
    print("pipelined on a non-default stream n=%d: %.3f ms/frame" % (n, e0.elapsed_time(e1) / n))
r.close()
