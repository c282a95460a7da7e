"""Multi-GPU sharding of the hot path (new design; the reference is single-GPU).

One process per GPU, `torch.distributed` for the plumbing (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  Work is partitioned only where it shards naturally:

* sort-first   a large frame is cut into a grid of rectangles (multiples of 8 px, each <= 2048 px --
               the reference's viewport limit, cuda/Constants.hpp:21); rank r renders the rectangles
               ``rects[r::world]`` through ``CudaRaster.setSubViewport`` with the geometry replicated.
* view-parallel independent views (cube-map faces, shadow cascades) go round-robin over the ranks.

The only exchange step is the composite: disjoint framebuffer rectangles (or whole frames) are
gathered to the display rank.  No reduction operator is needed because regions are disjoint.
"""
import math
import os

MAX_VIEWPORT = 2048


def grid_for(parts):
    """(columns, rows) of the sort-first grid for `parts` rectangles: 1x1, 2x1, 2x2, 4x2, 4x4, ..."""
    if parts < 1 or parts & (parts - 1):
        raise ValueError("the number of sort-first parts must be a power of two")
    lg = parts.bit_length() - 1
    return 1 << ((lg + 1) // 2), 1 << (lg // 2)


def parent_cells(full_w, full_h):
    """The parent viewports of a frame: the cells of the even grid of <= 2048 px cells that csrc/Context.cu prepareFrame()
    sets the triangles up in (ceil(full / 2048) cells per axis, cell size rounded up to 8 px).  Returns
    (cellW, cellH, columns, rows); every sort-first window must lie inside ONE of these cells."""
    ncx, ncy = -(-full_w // MAX_VIEWPORT), -(-full_h // MAX_VIEWPORT)
    return _cell(full_w, ncx), _cell(full_h, ncy), ncx, ncy


def min_parts(full_w, full_h):
    """Fewest rectangles a frame can be rendered in: one per parent cell."""
    _, _, ncx, ncy = parent_cells(full_w, full_h)
    return ncx * ncy


def _cell(full, n):
    return (-(-full // n) + 7) & ~7


def split_frame(full_w, full_h, parts):
    """Cuts the frame into AT LEAST `parts` rectangles (x0, y0, w, h), row-major order: first into the parent cells of
    prepareFrame() (see parent_cells), then every cell into the same power-of-two grid of sub-rectangles, so that no
    rectangle ever straddles a cell whatever the frame size (5K, 8K, odd widths).  Origins are multiples of 8 and the
    rectangles tile the frame exactly."""
    cw, ch, ncx, ncy = parent_cells(full_w, full_h)
    sub = 1
    while ncx * ncy * sub < parts:
        sub *= 2
    cols, rows = grid_for(sub)
    rects = []
    for cy in range(ncy):
        for cx in range(ncx):
            px0, py0 = cx * cw, cy * ch
            pw, ph = min(cw, full_w - px0), min(ch, full_h - py0)
            sw, sh = _cell(pw, cols), _cell(ph, rows)
            for r in range(rows):
                for c in range(cols):
                    x0, y0 = px0 + c * sw, py0 + r * sh
                    w, h = min(sw, px0 + pw - x0), min(sh, py0 + ph - y0)
                    if w > 0 and h > 0:
                        rects.append((x0, y0, w, h))
    rects.sort(key=lambda q: (q[1], q[0]))
    return rects


def rects_of_rank(rects, rank, world):
    """Rectangles rank `rank` renders (round robin keeps neighbours on different GPUs)."""
    return [(i, r) for i, r in enumerate(rects) if i % world == rank]


def views_of_rank(num_views, rank, world):
    """View indices rank `rank` renders."""
    return list(range(rank, num_views, world))


def gather_frames(frame, gather_list, dst=0):
    """Gathers one equally sized frame per rank to `dst` (NCCL gather over NVLink / gloo)."""
    import torch.distributed as dist
    dist.gather(frame, gather_list if dist.get_rank() == dst else None, dst=dst)


class AsyncFrameGather:
    """Gather of one frame per rank to `dst`, overlapped with the rendering of the next frame.

    The caller renders frame k into ``surfaces[k % depth]`` on its own stream and calls ``submit(k)``;
    the gather runs on a side stream (NCCL), and ``before_render(k)`` makes the render stream wait until
    the gather that last read ``surfaces[k % depth]`` is done.  ``finish()`` joins the side stream.
    On CPU tensors (gloo tests) everything degenerates to a synchronous gather."""

    def __init__(self, surfaces, world, rank, dst=0):
        import torch
        self.torch, self.surfaces, self.world, self.rank, self.dst = torch, list(surfaces), world, rank, dst
        self.cuda = self.surfaces[0].is_cuda
        self.recv = [[torch.empty_like(s) for _ in range(world)] if rank == dst else None for s in self.surfaces]
        if self.cuda:
            dev = self.surfaces[0].device
            self.side = torch.cuda.Stream(device=dev)
            self.rendered = [torch.cuda.Event() for _ in self.surfaces]
            self.gathered = [None for _ in self.surfaces]

    def before_render(self, k):
        i = k % len(self.surfaces)
        if self.cuda and self.gathered[i] is not None:
            self.torch.cuda.current_stream(self.surfaces[i].device).wait_event(self.gathered[i])
        return self.surfaces[i]

    def submit(self, k):
        import torch.distributed as dist
        i = k % len(self.surfaces)
        if not self.cuda:
            dist.gather(self.surfaces[i], self.recv[i], dst=self.dst)
            return
        torch = self.torch
        self.rendered[i].record(torch.cuda.current_stream(self.surfaces[i].device))
        self.side.wait_event(self.rendered[i])
        with torch.cuda.stream(self.side):
            dist.gather(self.surfaces[i], self.recv[i], dst=self.dst)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.gathered[i] = ev

    def finish(self):
        if self.cuda:
            self.torch.cuda.current_stream(self.surfaces[0].device).wait_stream(self.side)

    def frames(self, k):
        """The gathered frames of step k on `dst` (valid after finish() / a later before_render of the same slot)."""
        return self.recv[k % len(self.surfaces)]


class PeerMemoryUnavailable(RuntimeError):
    """Raised on EVERY rank when some rank cannot map the display rank's frame slots."""


class PeerFrameSink:
    """View-parallel / sort-first composite WITHOUT a gather: rank `dst` owns the frame slots of all ranks (`depth`
    deep) and exports them through CUDA IPC; every rank renders straight into its own slot, so the fine raster's colour
    stores cross NVLink / NVSwitch while the frame is still being rendered (crb200.h: crb_ipc_*).  A 32-bit mark per
    (slot, rank) says which frame the slot holds (`publish`, stream ordered, written after the frame).

        sink = PeerFrameSink(world, rank, frame_bytes, depth=2)
        raster.setSurfaces(sink.surface(k, (w, h)), depth_surface); raster.drawTriangles(asynchronous=True); sink.publish(k)
    """

    def __init__(self, world, rank, frame_bytes, depth=2, dst=0, device=None):
        import ctypes
        import torch
        import torch.distributed as dist
        from .binding import load_library
        self.lib, self.world, self.rank, self.depth, self.dst = load_library(), world, rank, depth, dst
        self.frame_bytes = (frame_bytes + 255) & ~255
        self.flag_ofs = self.frame_bytes * world * depth
        total = self.flag_ofs + 4 * world * depth
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        handle = torch.zeros(64, dtype=torch.uint8, device=dev)
        base = ctypes.c_void_p()
        ok = True
        if rank == dst:
            buf = ctypes.create_string_buffer(64)
            ok = self.lib.crb_ipc_alloc(total, ctypes.byref(base), buf) == 0
            if ok:
                handle.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
        dist.broadcast(handle, src=dst)
        if rank != dst:
            raw = bytes(handle.cpu().numpy().tobytes())
            ok = self.lib.crb_ipc_open(raw, ctypes.byref(base)) == 0 and not os.environ.get("CRB_TEST_FAIL_IPC")
        self.base = base.value if ok else None
        # every rank must agree: a box without peer access between some pair of GPUs fails on SOME ranks only
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0:
            if self.base:
                (self.lib.crb_ipc_free if rank == dst else self.lib.crb_ipc_close)(self.base)
            self.base = None
            raise PeerMemoryUnavailable("PeerFrameSink: CUDA IPC / peer access is not available between all GPUs of this job")
        dist.barrier()

    def slot_pointer(self, k, rank=None):
        r = self.rank if rank is None else rank
        return self.base + ((k % self.depth) * self.world + r) * self.frame_bytes

    def mark_pointer(self, k, rank=None):
        """Device address of the 32-bit frame mark of (slot k % depth, rank)."""
        r = self.rank if rank is None else rank
        return self.base + self.flag_ofs + 4 * ((k % self.depth) * self.world + r)

    def surface(self, k, size, num_samples=1):
        """The colour surface rank `self.rank` renders frame k into (memory of rank `dst`)."""
        from .binding import CudaSurface
        return CudaSurface.from_pointer(self.slot_pointer(k), size, CudaSurface.FORMAT_RGBA8, num_samples)

    def publish(self, k, stream=None):
        import ctypes
        import torch
        s = torch.cuda.current_stream().cuda_stream if stream is None else stream
        word = self.base + self.flag_ofs + 4 * ((k % self.depth) * self.world + self.rank)
        if self.lib.crb_ipc_signal(ctypes.c_void_p(word), k + 1, ctypes.c_void_p(s)) != 0:
            raise RuntimeError("PeerFrameSink: signal failed")

    # -- staged variant: render into LOCAL surfaces, let the DMA engines push finished frames into the slots ---------------
    def attach_local(self, surfaces):
        """surfaces: the local colour tensors frames are rendered into (frame k -> surfaces[k % len]); enables push()."""
        import torch
        self.local = list(surfaces)
        self.side = torch.cuda.Stream(device=self.local[0].device)
        self.rendered = [torch.cuda.Event() for _ in self.local]
        self.pushed = [None for _ in self.local]

    def before_render(self, k):
        """Makes the render stream wait until the copy that last read the local surface of frame k is done."""
        import torch
        i = k % len(self.local)
        if self.pushed[i] is not None:
            torch.cuda.current_stream(self.local[i].device).wait_event(self.pushed[i])
        return self.local[i]

    def push(self, k):
        """Frame k is rendered (on the current stream): copy it into this rank's slot on the side stream (DMA over NVLink,
        overlapped with the rendering of the next frame) and leave the frame mark."""
        import ctypes
        import torch
        i = k % len(self.local)
        t = self.local[i]
        self.rendered[i].record(torch.cuda.current_stream(t.device))
        self.side.wait_event(self.rendered[i])
        with torch.cuda.stream(self.side):
            if self.lib.crb_ipc_copy(ctypes.c_void_p(self.slot_pointer(k)), ctypes.c_void_p(t.data_ptr()), t.numel() * t.element_size(),
                                     ctypes.c_void_p(self.side.cuda_stream)) != 0:
                raise RuntimeError("PeerFrameSink: copy failed")
            self.publish(k, stream=self.side.cuda_stream)
            ev = torch.cuda.Event()
            ev.record(self.side)
        self.pushed[i] = ev

    def finish(self):
        import torch
        if getattr(self, "side", None) is not None:
            torch.cuda.current_stream(self.local[0].device).wait_stream(self.side)

    def read_marks(self, raster):
        """On `dst`: the frame number + 1 every (slot, rank) holds (blocking; raster = a CudaRaster of this process)."""
        import numpy as np
        return raster._dev_to_numpy(self.base + self.flag_ofs, 4 * self.world * self.depth, np.uint32).reshape(self.depth, self.world)

    def read_frame(self, raster, k, rank, nbytes):
        """On `dst`: a host copy (uint32) of the slot of `rank` for frame k (blocking)."""
        import numpy as np
        return raster._dev_to_numpy(self.slot_pointer(k, rank), nbytes, np.uint32)

    def close(self):
        import torch
        torch.cuda.synchronize()
        if self.rank == self.dst:
            self.lib.crb_ipc_free(self.base)
        else:
            self.lib.crb_ipc_close(self.base)


def composite_sort_first(local_rects, local_tiles, full_w, full_h, num_rects, dst=0, out=None):
    """Composites sort-first rectangles on rank `dst`.

    local_rects  [(index, (x0, y0, w, h))] this rank rendered, local_tiles the matching 2-D U32/int32
    tensors (rows >= h, cols >= w: rounded surfaces are accepted).  Every rank must own the same
    number of rectangles (pad with empty work otherwise).  Returns the [full_h, full_w] frame on
    `dst`, None elsewhere.  Rectangles are disjoint, so the composite is a pure gather + paste.
    """
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    per_rank = math.ceil(num_rects / world)
    assert len(local_rects) <= per_rank
    rects_all = [None] * world
    dist.all_gather_object(rects_all, [(i, tuple(r)) for i, r in local_rects])
    cw = max(r[2] for rr in rects_all for _, r in rr)
    ch = max(r[3] for rr in rects_all for _, r in rr)
    dev, dt = local_tiles[0].device, local_tiles[0].dtype
    send = torch.zeros((per_rank, ch, cw), dtype=dt, device=dev)
    for k, ((_, (x0, y0, w, h)), t) in enumerate(zip(local_rects, local_tiles)):
        send[k, :h, :w] = t[:h, :w]
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return None
    if out is None:
        out = torch.zeros((full_h, full_w), dtype=dt, device=dev)
    for r in range(world):
        for k, (_, (x0, y0, w, h)) in enumerate(rects_all[r]):
            out[y0:y0 + h, x0:x0 + w] = recv[r][k, :h, :w]
    return out


def render_sort_first(raster, crb, vb, ib, num_tris, full_w, full_h, pipe, rects, clear=((0.2, 0.4, 0.8, 1.0), 1.0), samples=1, device=None):
    """Renders this rank's rectangles of a full_w x full_h frame; returns [(index, rect)], [colour tensors], [depth tensors]."""
    out_c, out_d = [], []
    for _, (x0, y0, w, h) in rects:
        color = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8, samples, device=device or "cuda:%d" % raster.device)
        depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32, samples, device=device or "cuda:%d" % raster.device)
        raster.setSurfaces(color, depth)
        raster.setPixelPipe(None, pipe)
        raster.setVertexBuffer(vb, 0)
        raster.setIndexBuffer(ib, 0, num_tris)
        raster.setSubViewport(full_w, full_h, x0, y0)
        if clear is not None:
            raster.deferredClear(*clear)
        raster.drawTriangles()
        out_c.append(color.tensor)
        out_d.append(depth.tensor)
    raster.setSubViewport(0, 0, 0, 0)
    return rects, out_c, out_d
