"""CPU: the N > 1 host logic (cudaraster-linux_b200/multigpu.py) over torch.distributed `gloo`,
world_size 2: sort-first partition + composite, view-parallel assignment + gather.  The per-rank
rendering is done by the CPU oracle here (the GPU path renders the same windows with CUDA:
tests/test_gpu_parity.py::test_4k_sort_first_windows, bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import cudaraster_linux_b200 as crb
    from cudaraster_linux_b200 import multigpu
    from tests import util
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- sort-first: a 640x384 frame in 4 windows, 2 per rank, composited on rank 0
        fw, fh = 640, 384
        v, i = crb.scenes.random_soup(3000, seed=9, stride_floats=8, size=0.5)
        rects = multigpu.split_frame(fw, fh, 4)
        mine = multigpu.rects_of_rank(rects, rank, world)
        tiles_c, tiles_d = [], []
        for _, (x0, y0, w, h) in mine:
            g = util.draw_gold(v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0), threads=2)
            tiles_c.append(torch.from_numpy(g["color"].view(np.int32).copy()))
            tiles_d.append(torch.from_numpy(g["depth"].view(np.int32).copy()))
        comp_c = multigpu.composite_sort_first(mine, tiles_c, fw, fh, len(rects), dst=0)
        comp_d = multigpu.composite_sort_first(mine, tiles_d, fw, fh, len(rects), dst=0)
        # ---- view-parallel: 4 views round robin, frames gathered to rank 0
        views = crb.scenes.view_matrix_variants(4)
        frames = []
        for k in multigpu.views_of_rank(4, rank, world):
            g = util.draw_gold(crb.scenes.apply_view(v, views[k]), i, 256, 192, "gouraud", 3, threads=2)
            frames.append(torch.from_numpy(g["color"].view(np.int32).copy()))
        gathered = []
        for f in frames:
            lst = [torch.empty_like(f) for _ in range(world)] if rank == 0 else None
            multigpu.gather_frames(f, lst, dst=0)
            if rank == 0:
                gathered.append([t.numpy().view(np.uint32) for t in lst])
        # ---- the double-buffered gather bench.py uses (synchronous on CPU tensors): frame k of every rank lands in slot k % 2
        surf = [torch.zeros((4, 8), dtype=torch.int32) for _ in range(2)]
        ag = multigpu.AsyncFrameGather(surf, world, rank, dst=0)
        ag_ok = True
        for k in range(5):
            ag.before_render(k).fill_(100 * k + rank)
            ag.submit(k)
            ag.finish()
            if rank == 0:
                ag_ok = ag_ok and all(int(t[0, 0]) == 100 * k + r and bool((t == 100 * k + r).all()) for r, t in enumerate(ag.frames(k)))
        if rank == 0:
            full = util.draw_gold(v, i, fw, fh, "gouraud", 3, threads=2)
            ok = np.array_equal(comp_c.numpy().view(np.uint32), full["color"]) and np.array_equal(comp_d.numpy().view(np.uint32), full["depth"])
            for step, per_rank in enumerate(gathered):
                for r, frame in enumerate(per_rank):
                    k = multigpu.views_of_rank(4, r, world)[step]
                    ref = util.draw_gold(crb.scenes.apply_view(v, views[k]), i, 256, 192, "gouraud", 3, threads=2)
                    ok = ok and np.array_equal(frame, ref["color"])
            with open(out_path, "w") as fh_:
                fh_.write("ok" if (ok and ag_ok) else "mismatch")
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sort_first_and_view_parallel_world2(tmp_path):
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert open(out).read() == "ok"


def test_split_frame_properties():
    import cudaraster_linux_b200  # noqa: F401  (registers the package alias)
    from cudaraster_linux_b200 import multigpu
    from oracle.binding import parent_cell
    for fw, fh in ((3840, 2160), (1920, 1080), (2048, 2048), (1000, 600), (5120, 2880), (7680, 4320), (6000, 3000), (3010, 2000), (2049, 17)):
        for parts in (1, 2, 4, 8, 16):
            rects = multigpu.split_frame(fw, fh, parts)
            cover = np.zeros((fh, fw), np.int32)
            for (x0, y0, w, h) in rects:
                assert x0 % 8 == 0 and y0 % 8 == 0 and 0 < w <= 2048 and 0 < h <= 2048
                # the containment check of csrc/Context.cu prepareFrame(), replayed: a window lies inside ONE parent cell
                parent_cell(fw, x0, w)
                parent_cell(fh, y0, h)
                cover[y0:y0 + h, x0:x0 + w] += 1
            assert (cover == 1).all(), "rectangles must tile the frame exactly"
            assert len(rects) >= max(parts, multigpu.min_parts(fw, fh))
    assert multigpu.min_parts(3840, 2160) == 4 and multigpu.min_parts(1920, 1080) == 1
    assert multigpu.views_of_rank(48, 3, 8) == list(range(3, 48, 8))
