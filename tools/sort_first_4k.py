#!/usr/bin/env python
"""BASELINE config 5 (i): a 3840x2160 frame of a 4 M-triangle grid, sort-first split over the GPUs of one box.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sort_first_4k.py [--frames K]
    python tools/sort_first_4k.py                     # one GPU renders all rectangles

The frame is cut into max(N, 4) rectangles (a 4K frame exceeds the 2048 px viewport limit: at least 2x2); rank r renders
rectangles r, r+N, ... of every frame STRAIGHT into the full frame that lives in rank 0's memory (multigpu.PeerFrameSink +
crb_set_color_pitch): geometry replicated, no gather, no paste.  Rank 0 then compares the frame with the same rectangles
rendered by itself into a local image (bit-exact) and prints one JSON line (device-timed, max over ranks).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--composite", default="push", choices=["inplace", "push"],
                    help="inplace: the fine raster stores straight into the full frame in rank 0's memory (32-byte rows over NVLink); push: rectangles are rendered into "
                         "local surfaces and pasted into the full frame by the DMA engines on a side stream (2-D copy, overlapped with the next frame)")
    ap.add_argument("--no-bounds", action="store_true", help="without the per-chunk clip-space bounds (every rank then sets up every triangle)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import cudaraster_linux_b200 as crb
    from cudaraster_linux_b200 import multigpu

    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    fw, fh = 3840, 2160
    verts, idx = crb.scenes.grid_gouraud(2000, 1000, seed=0xC0DE0005)
    n_tris = idx.shape[0]
    rects = multigpu.split_frame(fw, fh, max(world, 4))
    mine = multigpu.rects_of_rank(rects, rank, world)
    vb, ib = torch.from_numpy(verts).to(dev), torch.from_numpy(idx).to(dev)
    raster = crb.CudaRaster(local)
    pipe = crb.pipe_name("gouraud", 0, 3)
    depths = [crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32, device=dev) for _, (x0, y0, w, h) in mine]
    sink = multigpu.PeerFrameSink(world, rank, fw * fh * 4, depth=2, device=dev) if world > 1 else None
    local_frames = [torch.zeros((fh, fw), dtype=torch.int32, device=dev) for _ in range(2)]
    # once per mesh: clip-space bounds per chunk of 256 triangles; a rank then skips the chunks outside its rectangle
    bounds = None if args.no_bounds else raster.computeChunkBounds(vb, ib, n_tris, verts.shape[1] * 4)

    def frame_base(k):
        return sink.slot_pointer(k, rank=0) if sink else local_frames[k % 2].data_ptr()

    push = args.composite == "push" and world > 1
    lib = crb.load_library()
    side = torch.cuda.Stream(device=dev) if push else None
    local_colors = [[crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8, device=dev) for _, (x0, y0, w, h) in mine] for _ in range(2)] if push else None
    pushed = [None, None]

    def render(k, base, my_rects, my_depths, asynchronous=True):
        import ctypes
        stream = torch.cuda.current_stream(dev)
        if push and pushed[k % 2] is not None:
            stream.wait_event(pushed[k % 2])          # the paste that last read this set of local surfaces
        for j, ((_, (x0, y0, w, h)), d) in enumerate(zip(my_rects, my_depths)):
            if push:
                raster.setSurfaces(local_colors[k % 2][j], d)
                raster.setColorPitch(0)
            else:
                raster.setSurfaces(crb.CudaSurface.from_pointer(base + 4 * (y0 * fw + x0), (w, h), crb.CudaSurface.FORMAT_RGBA8), d)
                raster.setColorPitch(fw)
            raster.setPixelPipe(None, pipe)
            raster.setVertexBuffer(vb, 0)
            raster.setIndexBuffer(ib, 0, n_tris)
            raster.setChunkBounds(bounds)
            raster.setSubViewport(fw, fh, x0, y0)
            raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
            raster.drawTriangles(asynchronous=asynchronous)
        if push:
            ev = torch.cuda.Event()
            ev.record(stream)
            side.wait_event(ev)
            for j, (_, (x0, y0, w, h)) in enumerate(my_rects):
                c = local_colors[k % 2][j]
                assert lib.crb_ipc_copy_2d(ctypes.c_void_p(base + 4 * (y0 * fw + x0)), fw * 4, ctypes.c_void_p(c.tensor.data_ptr()), c.rounded_size[0] * 4, w * 4, h,
                                           ctypes.c_void_p(side.cuda_stream)) == 0
            sink.publish(k, stream=side.cuda_stream)
            pushed[k % 2] = torch.cuda.Event()
            pushed[k % 2].record(side)
        elif sink:
            sink.publish(k)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for k in range(args.warmup):
        render(k, frame_base(k), mine, depths, asynchronous=False)
    sync_all()
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    if push:
        lib.crb_ipc_delay(__import__("ctypes").c_void_p(stream.cuda_stream), int(rank * 100000 / world))   # ranks out of phase: their pastes interleave at rank 0
    for k in range(args.frames):
        render(k, frame_base(k), mine, depths)
    if push:
        stream.wait_stream(side)
    e1.record(stream)
    raster.finish()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1) / args.frames], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())

    ok = None
    if rank == 0:
        # the same frame rendered by rank 0 alone into a local image
        k_last = args.frames - 1
        ref = torch.zeros((fh, fw), dtype=torch.int32, device=dev)
        all_depths = [crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32, device=dev) for (x0, y0, w, h) in rects]
        for (i, r), d in zip(list(enumerate(rects)), all_depths):
            raster.setSurfaces(crb.CudaSurface.from_pointer(ref.data_ptr() + 4 * (r[1] * fw + r[0]), (r[2], r[3]), crb.CudaSurface.FORMAT_RGBA8), d)
            raster.setColorPitch(fw)
            raster.setChunkBounds(None)                 # the check frame is rendered WITHOUT the chunk cull
            raster.setSubViewport(fw, fh, r[0], r[1])
            raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
            raster.drawTriangles()
        torch.cuda.synchronize(dev)
        got = sink.read_frame(raster, k_last, 0, fw * fh * 4) if sink else local_frames[k_last % 2].cpu().numpy().view(np.uint32).reshape(-1)
        want = ref.cpu().numpy().view(np.uint32).reshape(-1)
        ok = bool(np.array_equal(got, want)) and bool((want != want[0]).any())
        line = {"metric": "Mtris/s", "value": n_tris / (ms * 1e-3) / 1e6, "unit": "Mtris/s", "frames_per_s": 1e3 / ms, "ms_per_frame": ms, "n_gpus": world,
                "config": {"workload": "C5(i): 4M-triangle grid, Gouraud, depth test, 3840x2160, sort-first over %d rectangles" % len(rects),
                           "composite": ("rectangles rendered locally and pasted into rank 0's full frame by the DMA engines (2-D copy over NVLink on a side stream, overlapped with the next frame)" if push else
                                         "rectangles rendered in place into rank 0's full frame (CUDA IPC peer memory, crb_set_color_pitch)") if world > 1 else "single GPU, rectangles rendered in place",
                           "geometry": "replicated; " + ("every rank sets up all triangles" if args.no_bounds else "per-chunk (256 triangles) clip-space bounds, computed once per mesh: a rank skips the chunks outside its rectangle"),
                           "frame_equals_single_gpu_render": ok}, "scaling": "strong", "frames": args.frames}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    sync_all()
    if sink:
        sink.close()
    raster.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
