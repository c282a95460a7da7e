"""Shared helpers of the parity tests."""
import numpy as np
import torch

from oracle import binding as G

FLAG_DEPTH, FLAG_LERP = 1, 2
STRIDE = {"passthrough": 16, "gouraud": 32, "gouraudDiscard": 32, "gouraudQuads": 32, "gouraudCounters": 32, "gouraudTimers": 32, "texPhong": 64}


def draw_cuda(raster, crb, verts, idx, width, height, shader, flags, samples_log2=0, blend="BlendReplace", clear=(0.2, 0.4, 0.8, 1.0),
              clear_depth=1.0, init=None, sub=None, pipe=None):
    """Renders with the CUDA pipeline through the C ABI; returns (color, depth) as uint32 arrays."""
    n = 1 << samples_log2
    color = crb.CudaSurface((width, height), crb.CudaSurface.FORMAT_RGBA8, n)
    depth = crb.CudaSurface((width, height), crb.CudaSurface.FORMAT_DEPTH32, n)
    if init is not None:
        color.tensor.copy_(torch.from_numpy(init[0].view(np.int32)))
        depth.tensor.copy_(torch.from_numpy(init[1].view(np.int32)))
    vb = torch.from_numpy(np.ascontiguousarray(verts, np.float32)).cuda()
    ib = torch.from_numpy(np.ascontiguousarray(idx, np.int32)).cuda()
    raster.setSurfaces(color, depth)
    raster.setPixelPipe(None, pipe or crb.pipe_name(shader, samples_log2, flags, blend))
    raster.setVertexBuffer(vb, 0)
    raster.setIndexBuffer(ib, 0, idx.shape[0])
    if sub is None:
        raster.setSubViewport(0, 0, 0, 0)
    else:
        raster.setSubViewport(*sub)
    if clear is not None:
        raster.deferredClear(clear, clear_depth)
    raster.drawTriangles()
    torch.cuda.synchronize()
    return color.numpy(), depth.numpy()


def draw_gold(verts, idx, width, height, shader, flags, samples_log2=0, blend="BlendReplace", clear=(0.2, 0.4, 0.8, 1.0), clear_depth=1.0,
              init=None, sub=None, threads=8, want_counts=False):
    cv = None if clear is None else G.clear_values(clear, clear_depth)
    cfg = G.make_config(width, height, samples_log2, flags, STRIDE[shader], shader, blend, clear=cv, threads=threads, sub=sub)
    color = depth = None
    if init is not None:
        color, depth = init[0].copy(), init[1].copy()
    r = G.render(cfg, verts, idx, color, depth, want_counts=want_counts)
    return r


def gold_setup(verts, idx, width, height, shader, flags, samples_log2=0, sub=None):
    cfg = G.make_config(width, height, samples_log2, flags, STRIDE[shader], shader, "BlendReplace", sub=sub)
    return G.triangle_setup(cfg, verts, idx)


def defined_data_mask(flags):
    """Words of CRTriangleData that setup defines for these render-mode flags."""
    m = np.zeros(16, bool)
    if flags & FLAG_DEPTH:
        m[0:4] = True
    if flags & FLAG_LERP:
        m[4:13] = True
    m[12:16] = True  # vb (0 without lerp), vi0..vi2
    return m


def compare_setup(cuda_wb, gold_out, num_tris, flags):
    """Bit-exact comparison of triSubtris / triHeader / triData through the sub-triangle indirection."""
    cs, gs = cuda_wb["triSubtris"], gold_out["triSubtris"]
    assert np.array_equal(cs, gs), "triSubtris differ at %s" % np.nonzero(cs != gs)[0][:10]
    mask = defined_data_mask(flags)
    ch, cd, gh, gd = cuda_wb["triHeader"], cuda_wb["triData"], gold_out["triHeader"], gold_out["triData"]
    single = np.nonzero(cs == 1)[0]
    assert np.array_equal(ch[single], gh[single]), "single-triangle headers differ"
    assert np.array_equal(cd[single][:, mask], gd[single][:, mask]), "single-triangle data differ"
    multi = np.nonzero(cs > 1)[0]
    for t in multi:
        cb, gb, n = int(ch[t, 3]), int(gh[t, 3]), int(cs[t])
        assert np.array_equal(ch[cb:cb + n], gh[gb:gb + n]), "clipped triangle %d headers differ" % t
        assert np.array_equal(cd[cb:cb + n][:, mask], gd[gb:gb + n][:, mask]), "clipped triangle %d data differ" % t
    assert cuda_wb["counters"]["numSubtris"] == gold_out["numSubtris"]
    return len(single), len(multi)


def color_max_diff(a, b):
    d = 0
    for s in (0, 8, 16, 24):
        d = max(d, int(np.abs(((a >> s) & 0xFF).astype(np.int32) - ((b >> s) & 0xFF).astype(np.int32)).max()))
    return d
