#!/usr/bin/env python
"""BASELINE config 5 (ii): 48 views (6 cube faces x 8 positions) of the C2 mesh at 1024x1024, sharded round robin over the GPUs
of one box, every frame delivered into its slot in rank 0's memory.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/views_48.py [--reps K]
    python tools/views_48.py                          # one GPU renders all 48 views

Rank r renders views r, r+N, ...; all of a rank's frames are enqueued by ONE C call (crb_draw_batch_async): each frame is
rendered into one of two local surfaces and pushed into its slot on rank 0 by the DMA engines on the library's side stream,
overlapped with the next view (CUDA IPC peer memory over NVLink; N = 1: plain device copies).  Rank 0 re-renders six of the 48
views itself afterwards and compares them with what arrived (bit-exact).  One JSON line: device-timed, max over ranks, strong
scaling (the batch is 48 views whatever N is).
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5, help="the 48-view batch is rendered this many times (median)")
    ap.add_argument("--views", type=int, default=48)
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

    w = h = 1024
    nv = args.views
    verts, idx = crb.scenes.grid_gouraud(1000, 500)
    n_tris = idx.shape[0]
    views = crb.scenes.view_matrix_variants(nv)
    mine = multigpu.views_of_rank(nv, rank, world)
    ib = torch.from_numpy(idx).to(dev)
    vbs = {v: torch.from_numpy(crb.scenes.apply_view(verts, views[v])).to(dev) for v in mine}
    raster = crb.CudaRaster(local)
    raster.setPixelPipe(None, crb.pipe_name("gouraud", 0, 3))
    colors = [crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8, device=dev) for _ in range(2)]
    depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32, device=dev)
    frame_bytes = w * h * 4
    # rank 0 owns one slot per VIEW (+ one mark each); the others map them
    lib = crb.load_library()
    base = ctypes.c_void_p()
    handle = torch.zeros(64, dtype=torch.uint8, device=dev)
    total = frame_bytes * nv + 4 * nv
    if rank == 0:
        buf = ctypes.create_string_buffer(64)
        assert lib.crb_ipc_alloc(total, ctypes.byref(base), buf) == 0
        handle.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
    if world > 1:
        dist.broadcast(handle, src=0)
        if rank != 0:
            assert lib.crb_ipc_open(bytes(handle.cpu().numpy().tobytes()), ctypes.byref(base)) == 0, "CUDA IPC / peer access unavailable"
        dist.barrier()
    base = base.value

    def batch(rep):
        frames = []
        for j, v in enumerate(mine):
            frames.append({"color": colors[j % 2], "depth": depth, "vb": vbs[v], "ib": ib, "num_tris": n_tris, "clear": ((0.2, 0.4, 0.8, 1.0), 1.0),
                           "push_dst": base + v * frame_bytes, "push_bytes": frame_bytes, "slot": j % 2,
                           "signal_word": base + nv * frame_bytes + 4 * v, "signal_value": rep + 1})
        return raster.makeBatch(frames)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # sizes the work buffers (synchronous draw with overflow retry)
    raster.setSurfaces(colors[0], depth)
    raster.setVertexBuffer(vbs[mine[0]], 0)
    raster.setIndexBuffer(ib, 0, n_tris)
    raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
    raster.drawTriangles()
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for rep in range(args.reps + 1):
        b = batch(rep)
        sync_all()
        e0.record(stream)
        raster.drawBatch(b)
        raster.batchJoin()
        e1.record(stream)
        raster.finish()
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rep > 0:
            times.append(float(ms.item()))
    ms = float(np.median(times))

    if rank == 0:
        marks = raster._dev_to_numpy(base + nv * frame_bytes, 4 * nv, np.uint32)
        ok = bool((marks == args.reps + 1).all())
        for v in (0, 7, 13, 29, 38, nv - 1):
            vb = torch.from_numpy(crb.scenes.apply_view(verts, views[v])).to(dev)
            raster.setSurfaces(colors[0], depth)
            raster.setVertexBuffer(vb, 0)
            raster.setIndexBuffer(ib, 0, n_tris)
            raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
            raster.drawTriangles()
            torch.cuda.synchronize(dev)
            got = raster._dev_to_numpy(base + v * frame_bytes, frame_bytes, np.uint32)
            ok = ok and bool(np.array_equal(got, colors[0].numpy().reshape(-1)))
        line = {"metric": "Mtris/s", "value": nv * n_tris / (ms * 1e-3) / 1e6, "unit": "Mtris/s", "views_per_s": nv * 1e3 / ms, "ms_per_batch": ms, "n_gpus": world, "scaling": "strong",
                "config": {"workload": "C5(ii): %d views (6 cube faces x 8 positions) of the 1M-triangle C2 mesh, Gouraud, depth test, 1024x1024, round robin over the GPUs" % nv,
                           "composite": "every frame pushed into its slot in rank 0's memory by the DMA engines (library side stream, overlapped with the next view); one C call per rank for its whole share",
                           "frames_verified_on_rank0": ok}, "reps": args.reps}
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    sync_all()
    if rank == 0:
        lib.crb_ipc_free(ctypes.c_void_p(base))
    else:
        lib.crb_ipc_close(ctypes.c_void_p(base))
    raster.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
