"""GPU: the CUDA pipeline (through the C ABI) against the CPU oracle on the same seeded inputs.
Bit-exact for triSubtris / triHeader / triData, depth and colour; queues must be ordered supersets."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _check_surfaces(cc, cd, g, lsb=0):
    assert np.array_equal(cd, g["depth"]), "depth differs at %d texels" % int((cd != g["depth"]).sum())
    if lsb == 0:
        assert np.array_equal(cc, g["color"]), "colour differs at %d texels" % int((cc != g["color"]).sum())
    else:
        assert util.color_max_diff(cc, g["color"]) <= lsb


def test_c1_cube_known_answers(raster, crb):
    """BASELINE config 1 at both resolutions: every texel is background or cube."""
    for w, h in ((1024, 768), (720, 480)):
        v, i = crb.scenes.cube(w, h)
        cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "passthrough", 1, pipe="PixelPipe_passthrough")
        g = util.draw_gold(v, i, w, h, "passthrough", 1)
        _check_surfaces(cc, cd, g)
        assert set(np.unique(cc).tolist()) == {0xFF0000FF, 0xFFCC6633}
        assert cd.max() == 0xFFFFBB3F and (cd[cc == 0xFF0000FF] < 0xFFFFBB3F).all() and (cd[cc == 0xFFCC6633] == 0xFFFFBB3F).all()
        st = raster.getStats()
        assert all(st[k] >= 0 for k in st) and "triangleSetup" in raster.getProfilingInfo()


@pytest.mark.parametrize("shader,flags", [("passthrough", 1), ("passthrough", 0), ("gouraud", 3), ("gouraud", 2), ("gouraud", 1), ("gouraud", 0)])
def test_setup_and_surfaces_random_soup(raster, crb, shader, flags):
    """Mixed triangle soup with frustum-crossing and w<=0 triangles: setup records bit-exact
    (clipped ones through the misc indirection), then surfaces bit-exact."""
    w, h = 640, 360
    v, i = crb.scenes.random_soup(20000, seed=1234 + flags, stride_floats=util.STRIDE[shader] // 4)
    g = util.draw_gold(v, i, w, h, shader, flags)
    gs = util.gold_setup(v, i, w, h, shader, flags)
    try:
        # 0 = ordered sort, 3 = direct path with every record written in full, 1 = default (micro-triangles leave no header / depth row)
        for mode in (0, 3, 1):
            raster.setBinningMode(mode)
            cc, cd = util.draw_cuda(raster, crb, v, i, w, h, shader, flags)
            if mode != 1:
                wb = raster.getWorkBuffers(i.shape[0])
                n_single, n_multi = util.compare_setup(wb, gs, i.shape[0], flags)
                assert n_single > 1000 and n_multi > 100
            _check_surfaces(cc, cd, g)
    finally:
        raster.setBinningMode(1)


def test_queues_are_ordered_supersets(raster, crb, gold):
    """Bin and tile queues: per-bin / per-tile entries strictly increasing in triIdx*8+sub, and every
    (triangle, tile) pair with coverage is present."""
    import ctypes
    w, h = 640, 360
    v, i = crb.scenes.random_soup(6000, seed=77, stride_floats=4, size=0.6)
    raster.setBinningMode(0)   # the ordered two-level sort is what this test inspects
    util.draw_cuda(raster, crb, v, i, w, h, "passthrough", 1)
    raster.setBinningMode(1)
    wb = raster.getWorkBuffers(i.shape[0])
    tq, ts, tc = wb["tileQueue"], wb["tileStart"], wb["tileCount"]
    bq, bs, bt = wb["binQueue"], wb["binStart"], wb["binTotal"]
    for b in range(len(bs)):
        e = bq[bs[b]:bs[b] + bt[b]]
        assert (np.diff(e) > 0).all(), "bin %d not strictly increasing" % b
    tiles_x = (w + 7) // 8
    present = {}
    for t in range(len(ts)):
        e = tq[ts[t]:ts[t] + tc[t]]
        assert (np.diff(e) > 0).all(), "tile %d not strictly increasing" % t
        present[t] = set(e.tolist())
    # exact coverage from the oracle for a sample of sub-triangles
    L = gold.lib()
    hdr, sub = wb["triHeader"], wb["triSubtris"]
    rng = np.random.default_rng(5)
    checked = 0
    for tri in rng.permutation(i.shape[0])[:1500]:
        n = int(sub[tri])
        for s in range(n):
            di = tri if n == 1 else int(hdr[tri, 3]) + s
            entry = tri * 8 + (7 if n == 1 else s)
            hraw = np.ascontiguousarray(hdr[di])
            xs = hraw[:3].view(np.int16)[0::2].astype(np.int64) + w * 8
            ys = hraw[:3].view(np.int16)[1::2].astype(np.int64) + h * 8
            for ty in range(max(int(ys.min()) >> 7, 0), min(int(ys.max()) >> 7, (h + 7) // 8 - 1) + 1):
                for tx in range(max(int(xs.min()) >> 7, 0), min(int(xs.max()) >> 7, tiles_x - 1) + 1):
                    if L.gold_cover_tile(hraw.ctypes.data, w, h, tx, ty):
                        assert entry in present[tx + ty * tiles_x], "covered pair missing from tile queue"
                        checked += 1
    assert checked > 1000


@pytest.mark.parametrize("samples_log2", [1, 2, 3])
def test_msaa_gouraud(raster, crb, samples_log2):
    w, h = 320, 200
    v, i = crb.scenes.random_soup(8000, seed=99 + samples_log2, stride_floats=8)
    cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3, samples_log2)
    wb = raster.getWorkBuffers(i.shape[0])
    util.compare_setup(wb, util.gold_setup(v, i, w, h, "gouraud", 3, samples_log2), i.shape[0], 3)
    _check_surfaces(cc, cd, util.draw_gold(v, i, w, h, "gouraud", 3, samples_log2), lsb=1)


@pytest.mark.parametrize("shader,flags,samples_log2,blend", [
    ("gouraud", 3, 0, "BlendSrcOver"), ("gouraud", 2, 0, "BlendSrcOver"), ("gouraud", 3, 0, "BlendAdditive"), ("gouraud", 3, 2, "BlendSrcOver"),
    ("gouraud", 2, 2, "BlendSrcOver"), ("passthrough", 1, 0, "BlendDepthOnly"), ("gouraudDiscard", 3, 0, "BlendReplace"),
    ("gouraudDiscard", 3, 2, "BlendReplace"), ("texPhong", 3, 0, "BlendReplace"), ("texPhong", 3, 2, "BlendReplace"), ("passthrough", 1, 2, "BlendReplace")])
def test_blend_and_shader_variants(raster, crb, shader, flags, samples_log2, blend):
    """Ordered blending, discard and the Phong pipe: submission order must be honoured exactly."""
    w, h = 256, 192
    v, i = crb.scenes.random_soup(5000, seed=4242, stride_floats=util.STRIDE[shader] // 4, size=0.5)
    if shader.startswith("gouraud"):
        v[:, 7] = np.random.default_rng(3).uniform(0.2, 1.0, v.shape[0]).astype(np.float32)  # alpha
    cc, cd = util.draw_cuda(raster, crb, v, i, w, h, shader, flags, samples_log2, blend)
    g = util.draw_gold(v, i, w, h, shader, flags, samples_log2, blend)
    _check_surfaces(cc, cd, g, lsb=0 if shader == "passthrough" else 1)


def test_no_deferred_clear_accumulates(raster, crb):
    """Without a deferred clear the previous surface contents persist and untouched tiles stay untouched."""
    w, h = 320, 200
    rng = np.random.default_rng(11)
    rw, rh = (w + 7) & ~7, (h + 7) & ~7
    init = (rng.integers(0, 2**32, (rh, rw), dtype=np.uint32), rng.integers(2**31, 2**32, (rh, rw), dtype=np.uint32))
    v, i = crb.scenes.random_soup(300, seed=8, stride_floats=8, size=0.15)
    cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3, clear=None, init=init)
    g = util.draw_gold(v, i, w, h, "gouraud", 3, clear=None, init=init)
    _check_surfaces(cc, cd, g, lsb=1)
    assert (cc == init[0]).any() and (cc != init[0]).any()


def test_odd_viewport_and_empty_draw(raster, crb):
    """Viewport not a multiple of 8 (rounded surface), and a draw with zero triangles that only clears."""
    w, h = 333, 257
    v, i = crb.scenes.random_soup(4000, seed=21, stride_floats=8, size=0.4)
    cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3)
    _check_surfaces(cc, cd, util.draw_gold(v, i, w, h, "gouraud", 3), lsb=1)
    cc, cd = util.draw_cuda(raster, crb, v, i[:0], w, h, "gouraud", 3)
    assert (cc == 0xFFCC6633).all() and (cd == 0xFFFFBB3F).all()


def test_sort_first_window_matches_full_frame(raster, crb):
    """A frame rendered as four sub-viewports equals the same frame rendered whole (CUDA and oracle)."""
    fw, fh = 512, 384
    v, i = crb.scenes.random_soup(6000, seed=31, stride_floats=8, size=0.5)
    full_c, full_d = util.draw_cuda(raster, crb, v, i, fw, fh, "gouraud", 3)
    _check_surfaces(full_c, full_d, util.draw_gold(v, i, fw, fh, "gouraud", 3), lsb=1)
    for (x0, y0, w, h) in [(0, 0, 256, 192), (256, 0, 256, 192), (0, 192, 256, 192), (256, 192, 256, 192)]:
        cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0))
        g = util.draw_gold(v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0))
        _check_surfaces(cc, cd, g, lsb=1)
        assert np.array_equal(cd, full_d[y0:y0 + h, x0:x0 + w])
        assert np.array_equal(cc, full_c[y0:y0 + h, x0:x0 + w])
    raster.setSubViewport(0, 0, 0, 0)


def test_c2_reduced_and_host_entry(raster, crb):
    """C2-style grid (reduced to 100k triangles) bit-exact, and the host-buffer entry returns the same frame."""
    import torch
    w, h = 1920, 1080
    v, i = crb.scenes.grid_gouraud(250, 200)
    cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3)
    g = util.draw_gold(v, i, w, h, "gouraud", 3)
    _check_surfaces(cc, cd, g, lsb=1)
    hv = torch.from_numpy(v).pin_memory()
    hi = torch.from_numpy(i).pin_memory()
    hc = torch.zeros((1080, 1920), dtype=torch.int32).pin_memory()
    hd = torch.zeros((1080, 1920), dtype=torch.int32).pin_memory()
    raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
    raster.drawTrianglesHost(hv, hi, i.shape[0], hc, hd)
    assert np.array_equal(hc.numpy().view(np.uint32), cc) and np.array_equal(hd.numpy().view(np.uint32), cd)


def test_c2_full_size_properties(raster, crb):
    """BASELINE config 2 at full size (1M triangles, 1080p): depth bit-exact against the oracle
    (8 host threads) and size-independent properties: idempotence, full coverage, counters."""
    w, h = 1920, 1080
    v, i = crb.scenes.grid_gouraud(1000, 500)
    cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3)
    cc2, cd2 = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3)
    assert np.array_equal(cc, cc2) and np.array_equal(cd, cd2)           # deterministic / idempotent
    assert (cd < 0xFFFFBB3F).all()                                        # the mesh covers the whole frame
    c = raster.getCounters()
    assert raster.lastFrameDirect()                                       # automatic mode: order-independent pipe -> direct path
    assert c["overflow"] == 0 and c["numLargeTris"] == 0 and c["numBinEntries"] == 0
    g = util.draw_gold(v, i, w, h, "gouraud", 3)
    _check_surfaces(cc, cd, g, lsb=1)


def test_async_frames_match_sync(raster, crb):
    """crb_draw_triangles_async + crb_finish: frames enqueued back to back equal the synchronous result;
    an overflowing asynchronous frame is reported by finish()."""
    import torch
    w, h = 640, 360
    v, i = crb.scenes.random_soup(20000, seed=5, stride_floats=8, size=0.3)
    cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3)           # synchronous: sizes the buffers
    for _ in range(3):
        raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
        raster.drawTriangles(asynchronous=True)
    raster.finish()
    torch.cuda.synchronize()
    color, depth = raster._keep["color"], raster._keep["depth"]
    assert np.array_equal(color.numpy(), cc) and np.array_equal(depth.numpy(), cd)
    assert raster.getCounters()["overflow"] == 0


def test_4k_sort_first_windows(raster, crb):
    """BASELINE config 5(i) shape: a 3840x2160 frame (beyond the 2048 px viewport limit) as 8 windows
    of 960x1080; every window bit-exact against the oracle, and the 2x2 split gives the same frame."""
    from cudaraster_linux_b200 import multigpu
    fw, fh = 3840, 2160
    v, i = crb.scenes.grid_gouraud(300, 200)
    frames = {}
    for parts in (8, 4):
        color = np.zeros((fh, fw), np.uint32)
        depth = np.zeros((fh, fw), np.uint32)
        for (x0, y0, w, h) in multigpu.split_frame(fw, fh, parts):
            cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0))
            if parts == 8:
                g = util.draw_gold(v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0))
                _check_surfaces(cc, cd, g, lsb=1)
            color[y0:y0 + h, x0:x0 + w] = cc[:h, :w]
            depth[y0:y0 + h, x0:x0 + w] = cd[:h, :w]
        frames[parts] = (color, depth)
    raster.setSubViewport(0, 0, 0, 0)
    assert np.array_equal(frames[8][0], frames[4][0]) and np.array_equal(frames[8][1], frames[4][1])
    assert (frames[8][1] < 0xFFFFBB3F).mean() > 0.9


def test_pipelined_host_frames(raster, crb):
    """crb_draw_triangles_host_async: six DIFFERENT frames in flight (upload / render / download
    overlapped, double-buffered staging), each landing in its own host buffer, each bit-exact."""
    import torch
    w, h = 640, 360
    scenes = [crb.scenes.random_soup(6000 + 1500 * k, seed=900 + k, stride_floats=8, size=0.25 + 0.05 * k) for k in range(6)]
    v0, i0 = scenes[-1]
    util.draw_cuda(raster, crb, v0, i0, w, h, "gouraud", 3)                 # synchronous: sizes the work buffers for the largest frame
    hv = [torch.from_numpy(v).pin_memory() for v, _ in scenes]
    hi = [torch.from_numpy(i).pin_memory() for _, i in scenes]
    hc = [torch.zeros((360, 640), dtype=torch.int32).pin_memory() for _ in scenes]
    hd = [torch.zeros((360, 640), dtype=torch.int32).pin_memory() for _ in scenes]
    for k, (v, i) in enumerate(scenes):
        raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
        raster.drawTrianglesHostAsync(hv[k], hi[k], i.shape[0], hc[k], hd[k])
    raster.finish()
    for k, (v, i) in enumerate(scenes):
        g = util.draw_gold(v, i, w, h, "gouraud", 3)
        _check_surfaces(hc[k].numpy().view(np.uint32), hd[k].numpy().view(np.uint32), g, lsb=1)


def test_back_to_back_frames_of_different_shape(raster, crb):
    """No memset between frames: the count matrices clean themselves and the counters are handed over
    by the fine raster kernel.  Alternate scenes / viewports / sample counts (layout changes, empty
    draws, a frame that overflows its queues and is retried) and compare every frame."""
    cases = [(640, 360, 20000, 0, 0.3), (320, 200, 3000, 0, 0.8), (640, 360, 20000, 2, 0.3), (1920, 1080, 60000, 0, 0.05), (640, 360, 0, 0, 0.3),
             (640, 360, 25000, 0, 1.5), (320, 200, 3000, 0, 0.8)]
    for rep in range(2):
        for w, h, n, s_log2, size in cases:
            # the 1080p case keeps its triangles with a vertex at w ~ 0-: their clipped remains can span the whole guard band with
            # an overflowed depth plane whose values fall BELOW the triangle's own zmin, which is why the tile-level early-Z
            # only drops what the plane itself proves invisible (earlyZCull, FineRaster.cuh) and the frame equals the oracle's
            v, i = crb.scenes.random_soup(max(n, 1), seed=4000 + n + rep, stride_floats=8, size=size)
            if n == 0:
                i = i[:0]
            cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3, s_log2)
            _check_surfaces(cc, cd, util.draw_gold(v, i, w, h, "gouraud", 3, s_log2), lsb=1)


@pytest.mark.parametrize("shader,flags,samples_log2,blend", [
    ("gouraudQuads", 7, 0, "BlendReplace"),    # visibility first, then every quad shades its distinct winners
    ("gouraudQuads", 7, 0, "BlendSrcOver"),    # in-order shading on converged quads, helper lanes included
    ("gouraudQuads", 6, 0, "BlendSrcOver"),    # no depth test
    ("gouraudQuads", 7, 2, "BlendReplace"),    # MSAA: helper pixels shaded at the centroid of their own mask
    ("gouraudQuads", 7, 1, "BlendSrcOver"),
    ("gouraudDiscard", 7, 0, "BlendReplace")])  # a discarding shader: discarded lanes keep helping
def test_quads_mode(raster, crb, shader, flags, samples_log2, blend):
    """RenderModeFlag_EnableQuads (reference: cuda/PixelPipe.hpp:34, :59-69, FineRaster.inl:396-430, :705-724,
    :1055-1090): dFdx / dFdy inside 2x2 quads, shader run on helper pixels, ROP only where covered."""
    w, h = 256, 192
    for seed, size in ((515, 0.5), (516, 0.08)):
        v, i = crb.scenes.random_soup(4000, seed=seed, stride_floats=8, size=size)
        v[:, 7] = np.random.default_rng(3).uniform(0.2, 1.0, v.shape[0]).astype(np.float32)  # alpha
        cc, cd = util.draw_cuda(raster, crb, v, i, w, h, shader, flags, samples_log2, blend)
        g = util.draw_gold(v, i, w, h, shader, flags, samples_log2, blend)
        _check_surfaces(cc, cd, g, lsb=1)
    if shader == "gouraudQuads":
        assert len(np.unique(cc & 0xFFFF)) > 50, "derivative channels look constant"


@pytest.mark.parametrize("samples_log2", [0, 1, 2, 3])
def test_msaa_resolve_matches_oracle(raster, crb, gold, samples_log2, tmp_path):
    """crb_resolve_surface (CudaSurface::resolveToScreen, CudaSurface.hpp:73): box filter over the tile-replicated
    sample layout, bit-exact against the oracle, odd sizes included; PPM writer round trip."""
    n = 1 << samples_log2
    for w, h in ((322, 201), (64, 8)):
        v, i = crb.scenes.random_soup(3000, seed=21 + samples_log2, stride_floats=8, size=0.4)
        color = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8, n)
        depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32, n)
        import torch
        vb, ib = torch.from_numpy(v).cuda(), torch.from_numpy(i).cuda()
        raster.setSurfaces(color, depth)
        raster.setPixelPipe(None, crb.pipe_name("gouraud", samples_log2, 3))
        raster.setVertexBuffer(vb, 0)
        raster.setIndexBuffer(ib, 0, i.shape[0])
        raster.setSubViewport(0, 0, 0, 0)
        raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
        raster.drawTriangles()
        for flip in (False, True):
            got = color.resolve(flip_y=flip).cpu().numpy().view(np.uint32)
            want = gold.resolve(color.numpy(), w, h, n, flip_y=flip)
            assert np.array_equal(got, want)
    path = str(tmp_path / "frame.ppm")
    color.writePPM(path)
    raw = open(path, "rb").read()
    head = b"P6\n%d %d\n255\n" % (w, h)
    assert raw.startswith(head) and len(raw) == len(head) + w * h * 3
    img = np.frombuffer(raw[len(head):], np.uint8).reshape(h, w, 3)
    want = gold.resolve(color.numpy(), w, h, n, flip_y=True)
    assert np.array_equal(img[..., 0], want & 0xFF) and np.array_equal(img[..., 2], (want >> 16) & 0xFF)


def test_vertex_shader_stage_then_draw(raster, crb, gold):
    """The demo's frame (test/SceneCR.cpp:263-297): vertex shader kernel -> drawTriangles on the same stream.
    Shaded vertices bit-exact against the oracle's restatement, then the frame against the oracle."""
    import torch
    rng = np.random.default_rng(5)
    nv, nt, w, h = 5000, 9000, 400, 300
    vin = np.zeros((nv, 7), np.float32)
    vin[:, 0:3] = rng.uniform(-1.5, 1.5, (nv, 3))
    vin[:, 3:7] = rng.uniform(0, 1, (nv, 4))
    idx = rng.integers(0, nv, (nt, 3)).astype(np.int32)
    idx[:, 1] = (idx[:, 0] + rng.integers(1, 40, nt)) % nv
    idx[:, 2] = (idx[:, 0] + rng.integers(1, 40, nt)) % nv
    m = crb.scenes.cube_mvp(w, h)            # column-major 4x4, perspective * lookAt of the demo camera
    d_in = torch.from_numpy(vin).cuda()
    d_out = torch.zeros((nv, 8), dtype=torch.float32, device="cuda")
    raster.launchVertexShader(None, "vertexShader_color", d_in, d_out, nv, np.ascontiguousarray(m, np.float32).tobytes())
    want = gold.vertex_shader(m, vin, 8)
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), want.view(np.uint32))
    cc, cd = util.draw_cuda(raster, crb, want, idx, w, h, "gouraud", 3)
    g = util.draw_gold(want, idx, w, h, "gouraud", 3)
    _check_surfaces(cc, cd, g, lsb=1)
    # pass-through variant: positions only (test/shader/PassThrough.cu:16-35)
    d_in2 = torch.from_numpy(np.ascontiguousarray(vin[:, 0:3])).cuda()
    d_out2 = torch.zeros((nv, 4), dtype=torch.float32, device="cuda")
    raster.launchVertexShader(None, "vertexShader_passthrough", d_in2, d_out2, nv, np.ascontiguousarray(m, np.float32).tobytes())
    assert np.array_equal(d_out2.cpu().numpy().view(np.uint32), gold.vertex_shader(m, vin[:, 0:3], 4).view(np.uint32))


# ---- direct tile path (crb_set_binning_mode): unordered tile queues + order-independent resolve -------------------
def _small_scenes(crb):
    rng = np.random.default_rng(9)
    v, i = crb.scenes.grid_gouraud(160, 100)
    yield "grid", v, i
    # the same mesh submitted twice with different colours: every fragment of the second copy ties in depth with
    # the first and must lose (strict LESS in submission order == (depth, index) minimum)
    v2 = np.concatenate([v, v]); v2[v.shape[0]:, 4:8] = rng.uniform(0, 1, (v.shape[0], 4)).astype(np.float32)
    i2 = np.concatenate([i, i + v.shape[0]])
    yield "duplicate", v2, i2
    # the second copy FIRST in memory order but the draw order interleaved: ties resolved by index, not by position
    perm = rng.permutation(i2.shape[0])
    yield "shuffled-duplicate", v2, np.ascontiguousarray(i2[perm])
    vs, js = crb.scenes.random_soup(30000, seed=41, stride_floats=8, size=0.03, clip_fraction=0.02, behind_fraction=0.0)
    yield "soup", vs, js


@pytest.mark.parametrize("shader,flags,samples_log2,blend", [
    ("gouraud", 3, 0, "BlendReplace"), ("gouraud", 1, 0, "BlendReplace"), ("passthrough", 1, 0, "BlendDepthOnly"), ("gouraud", 3, 2, "BlendReplace"),
    ("gouraud", 3, 3, "BlendReplace"), ("gouraudQuads", 7, 0, "BlendReplace")])
def test_direct_tile_path_matches_oracle(raster, crb, shader, flags, samples_log2, blend):
    w, h = 512, 384
    try:
        for name, v, i in _small_scenes(crb):
            if shader == "passthrough":
                v = np.ascontiguousarray(v[:, :4])
            g = util.draw_gold(v, i, w, h, shader, flags, samples_log2, blend)
            queued = None
            for mode in (3, 2, 0):   # direct with tile queues only, direct + micro-triangle buffer, general
                raster.setBinningMode(mode)
                cc, cd = util.draw_cuda(raster, crb, v, i, w, h, shader, flags, samples_log2, blend)
                assert raster.lastFrameDirect() == (mode != 0), "%s: wrong path (mode %d)" % (name, mode)
                c = raster.getCounters()
                assert c["overflow"] == 0 and (c["numBinEntries"] == 0) == (mode != 0)
                if mode == 3:
                    queued = c["numTileEntries"]
                    assert queued > 10000                                  # everything went through the tile queues
                elif mode == 2 and samples_log2 == 0:
                    assert c["numTileEntries"] < queued, name            # part of the triangles went the micro way instead
                elif mode == 2:
                    assert c["numTileEntries"] == queued                   # MSAA: no micro path
                _check_surfaces(cc, cd, g, lsb=0 if shader == "passthrough" else 1)
    finally:
        raster.setBinningMode(1)


def test_direct_tile_path_large_triangles_and_auto(raster, crb):
    """Large triangles on the direct path are counted and scattered by whole CTAs (rows of tiles shared out over the threads, the
    covered span of a row found by bisection); automatic mode takes the direct path on every frame of an order-independent pipe;
    pipes that read dst never go direct."""
    w, h = 512, 384
    v, i = crb.scenes.grid_gouraud(160, 100)
    big = np.array([[-0.9, -0.8, 0.5, 1, 1, 0, 0, 1], [0.9, -0.7, 0.5, 1, 0, 1, 0, 1], [0.1, 0.9, 0.5, 1, 0, 0, 1, 1],
                    [-3.0, -3.0, 0.7, 1, 1, 1, 0, 1], [3.0, -3.0, 0.7, 1, 0, 1, 1, 1], [0.0, 3.0, 0.7, 1, 1, 0, 1, 1]], np.float32)
    nv = v.shape[0]
    vb = np.concatenate([v, big]); ib = np.concatenate([i[:1000], np.array([[nv, nv + 1, nv + 2]], np.int32), i[1000:], np.array([[nv + 3, nv + 4, nv + 5]], np.int32)])
    vs, js = crb.scenes.random_soup(5000, seed=1234, stride_floats=8)   # big, clipped and w<=0 triangles
    try:
        for mode, vv, ii in ((2, vb, ib), (2, vs, js), (3, vb, ib), (3, vs, js)):
            raster.setBinningMode(mode)
            cc, cd = util.draw_cuda(raster, crb, vv, ii, w, h, "gouraud", 3)
            assert raster.lastFrameDirect() and raster.getCounters()["numLargeTris"] > 0
            _check_surfaces(cc, cd, util.draw_gold(vv, ii, w, h, "gouraud", 3), lsb=1)
        raster.setBinningMode(2)
        for vv, ii in ((vb, ib), (vs, js)):
            cc, cd = util.draw_cuda(raster, crb, vv, ii, w, h, "gouraud", 3, 2)
            assert raster.lastFrameDirect()
            _check_surfaces(cc, cd, util.draw_gold(vv, ii, w, h, "gouraud", 3, 2), lsb=1)
        cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3, blend="BlendSrcOver")   # reads dst: order matters
        assert not raster.lastFrameDirect()
        _check_surfaces(cc, cd, util.draw_gold(v, i, w, h, "gouraud", 3, blend="BlendSrcOver"), lsb=1)
        # automatic mode: the direct path on EVERY frame of an order-independent pipe -- the first frame of a shape, frames whose
        # triangle count changes, frames with large triangles inside (no previous-frame heuristic)
        raster.setBinningMode(1)
        for k, n in enumerate((i.shape[0], 1000, 31999, 12, i.shape[0] - 7, 0, 5000)):
            ii = i[:n]
            cc, cd = util.draw_cuda(raster, crb, v, ii, w, h, "gouraud", 3)
            assert raster.lastFrameDirect(), "frame %d (%d triangles) left the direct path" % (k, n)
            _check_surfaces(cc, cd, util.draw_gold(v, ii, w, h, "gouraud", 3), lsb=1)
        for vv, ii in ((vb, ib), (vs, js)):
            cc, cd = util.draw_cuda(raster, crb, vv, ii, w, h, "gouraud", 3)
            assert raster.lastFrameDirect() and raster.getCounters()["numLargeTris"] > 0
            _check_surfaces(cc, cd, util.draw_gold(vv, ii, w, h, "gouraud", 3), lsb=1)
        # more large sub-triangles than the initial capacity of the global large list (16 384): overflow -> grow -> rerun, like the queues
        vm, im = crb.scenes.random_soup(150000, seed=77, stride_floats=8, size=0.3, clip_fraction=0.0, behind_fraction=0.0)
        cc, cd = util.draw_cuda(raster, crb, vm, im, 1024, 768, "gouraud", 3)
        assert raster.lastFrameDirect() and raster.getCounters()["numLargeTris"] > 16384 and raster.getCounters()["overflow"] == 0
        _check_surfaces(cc, cd, util.draw_gold(vm, im, 1024, 768, "gouraud", 3), lsb=1)
        fresh = crb.CudaRaster(0)   # and the very first frame of a context
        try:
            cc, cd = util.draw_cuda(fresh, crb, vb, ib, w, h, "gouraud", 3)
            assert fresh.lastFrameDirect()
            _check_surfaces(cc, cd, util.draw_gold(vb, ib, w, h, "gouraud", 3), lsb=1)
        finally:
            fresh.close()
    finally:
        raster.setBinningMode(1)


def test_direct_tile_path_sort_first_windows(raster, crb):
    """Sort-first windows on the direct path (micro-triangles included): every window equals its part of the full
    frame rendered on the general path, for a mesh of small triangles with a few large ones crossing the seams."""
    fw, fh = 512, 384
    v, i = crb.scenes.grid_gouraud(200, 150)
    big = np.array([[-0.7, -0.9, 0.6, 1, 1, 0, 0, 1], [0.8, -0.2, 0.6, 1, 0, 1, 0, 1], [-0.1, 0.8, 0.6, 1, 0, 0, 1, 1]], np.float32)
    nv = v.shape[0]
    v = np.concatenate([v, big]); i = np.concatenate([i, np.array([[nv, nv + 1, nv + 2]], np.int32)])
    try:
        raster.setBinningMode(0)
        full_c, full_d = util.draw_cuda(raster, crb, v, i, fw, fh, "gouraud", 3)
        _check_surfaces(full_c, full_d, util.draw_gold(v, i, fw, fh, "gouraud", 3), lsb=1)
        for mode in (2, 3):
            raster.setBinningMode(mode)
            for (x0, y0, w, h) in [(0, 0, 256, 192), (256, 0, 256, 192), (0, 192, 256, 192), (256, 192, 256, 192), (128, 96, 200, 120)]:
                cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0))
                assert raster.lastFrameDirect()
                rh, rw = cd.shape
                assert np.array_equal(cd[:h, :w], full_d[y0:y0 + h, x0:x0 + w]) and np.array_equal(cc[:h, :w], full_c[y0:y0 + h, x0:x0 + w])
                _check_surfaces(cc, cd, util.draw_gold(v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0)), lsb=1)
    finally:
        raster.setSubViewport(0, 0, 0, 0)
        raster.setBinningMode(1)


def test_direct_tile_path_async_frames_and_no_clear(raster, crb):
    """Asynchronous frames on the direct path (counters self-clean between frames), then a frame without a deferred
    clear: fragments must beat the depth already in the surface, ties with it fail."""
    import torch
    w, h = 512, 384
    v, i = crb.scenes.grid_gouraud(160, 100)
    g = util.draw_gold(v, i, w, h, "gouraud", 3)
    try:
        raster.setBinningMode(2)
        cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3)
        color = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8, 1)
        depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32, 1)
        vbuf, ibuf = torch.from_numpy(v).cuda(), torch.from_numpy(i).cuda()
        raster.setSurfaces(color, depth)
        raster.setVertexBuffer(vbuf, 0)
        raster.setIndexBuffer(ibuf, 0, i.shape[0])
        for k in range(6):
            raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
            raster.drawTriangles(asynchronous=True)
        raster.finish()
        assert raster.lastFrameDirect()
        _check_surfaces(color.numpy(), depth.numpy(), g, lsb=1)
        # second pass over the finished frame without a clear: every fragment ties with the stored depth and fails
        raster.drawTriangles()
        assert raster.lastFrameDirect()
        _check_surfaces(color.numpy(), depth.numpy(), g, lsb=1)
        rng = np.random.default_rng(11)
        init = (rng.integers(0, 2**32, g["color"].shape, dtype=np.uint32), rng.integers(2**31, 2**32, g["depth"].shape, dtype=np.uint32))
        cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3, clear=None, init=init)
        assert raster.lastFrameDirect()
        _check_surfaces(cc, cd, util.draw_gold(v, i, w, h, "gouraud", 3, clear=None, init=init), lsb=1)
    finally:
        raster.setBinningMode(1)


@pytest.mark.parametrize("samples_log2", [0, 2])
def test_profiling_mode_counters(raster, crb, samples_log2):
    """A pipe compiled with CR_PROFILING_MODE = ProfilingMode_Counters (reference: cuda/PrivateDefs.hpp:161-205,
    CudaRaster.cpp:424-450): same frame as the plain pipe, and a counters report whose numbers agree with the oracle."""
    import re
    w, h = 320, 200
    v, i = crb.scenes.random_soup(8000, seed=77, stride_floats=8, size=0.3)
    cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraudCounters", 3, samples_log2)
    info = raster.getProfilingInfo()
    g = util.draw_gold(v, i, w, h, "gouraud", 3, samples_log2, want_counts=True)
    _check_surfaces(cc, cd, g, lsb=1)
    assert "ProfilingMode_Counters" in info and "TriangleSetup:" in info and "FineRaster:" in info
    val = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"- ([A-Za-z./ ]+?)\s+([0-9.]+)%?\n", info)}
    for k in ("Viewport cull", "Backface cull", "Between pixels cull", "Clipped", "Early Z kill", "Empty coverage", "Z kills"):
        assert 0.0 <= val[k] <= 100.0, (k, val)
    gs = util.gold_setup(v, i, w, h, "gouraud", 3, samples_log2)
    sub = gs["triSubtris"]
    # culled = viewport + backface + between-pixels culls of unclipped triangles + clipped triangles that vanished
    assert abs((val["Viewport cull"] + val["Backface cull"] + val["Between pixels cull"]) - 100.0 * ((sub == 0).sum() - 0) / len(sub)) < val["Clipped"] + 0.1
    assert val["Clipped"] > 5.0 and val["Avg. tri/tile"] > 1 and val["Avg. frag/tri"] > 1
    # bin / coarse stages (reference: cuda/PrivateDefs.hpp:168-187): every line of the reference's report, sane values
    assert "BinRaster:" in info and "CoarseRaster:" in info
    for k in ("Input overflows", "Avg. triangles/round", "Avg. tri bb size", "Coverage single path", "Coverage fast path", "Coverage slow path", "Segment allocs/round",
              "Bins", "Rounds / Bin", "Merge / Round", "Triangles / Round", "Tiles / Round", "Emits / Round", "Allocs / Round", "Emits / Triangle", "Case A", "Case B", "Case C"):
        assert k in val, k
    assert 0 < val["Avg. triangles/round"] <= 32 and val["Avg. tri bb size"] >= 1.0
    assert abs(val["Coverage single path"] + val["Coverage fast path"] + val["Coverage slow path"] - 100.0) < 0.5 and val["Coverage slow path"] > 0.0
    nbins = ((w + 127) // 128) * ((h + 127) // 128)
    assert 1 <= val["Bins"] <= nbins and val["Rounds / Bin"] >= 1 and 0 < val["Triangles / Round"] <= 32
    assert val["Emits / Triangle"] >= 0.5 and val["Tiles / Round"] >= 1 and val["Emits / Round"] >= val["Tiles / Round"]
    assert abs(val["Case A"] + val["Case C"] - 100.0) <= 1.0 and val["Case B"] == 0
    c2 = raster.getCounters()
    assert abs(val["Emits / Round"] * val["Rounds / Bin"] * val["Bins"] - c2["numTileEntries"]) <= 0.06 * c2["numTileEntries"] + 64      # the counters add up to the queue
    c = g["counts"]
    tiles = ((w + 7) // 8) * ((h + 7) // 8)
    if samples_log2 == 0:   # fragments the fine raster saw = all covered (triangle, pixel) pairs minus those of early-Z-culled triangles
        assert 0.3 * c["fragments"] <= val["Avg. frag/tile"] * tiles <= 1.001 * c["fragments"] + tiles
    # the plain pipe is unaffected
    cc2, cd2 = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3, samples_log2)
    assert np.array_equal(cc, cc2) and np.array_equal(cd, cd2)
    assert "ProfilingMode_Default" in raster.getProfilingInfo()


def test_limits_and_error_behaviour(crb):
    """The reference's state checks, in its order and with its messages (CudaRaster.cpp:141-159, :243-260;
    CudaSurface.cpp:44-62), and its format limits: 2048 x 2048 viewport with 8 samples, degenerate input."""
    import torch
    r = crb.CudaRaster(0)
    try:
        for size, n, msg in (((0, 10), 1, "Size must be positive"), ((4096, 16), 1, "CR_MAXVIEWPORT_SIZE exceeded"), ((64, 64), 16, "cannot exceed 8"),
                             ((64, 64), 3, "power of two")):
            with pytest.raises(crb.CrbError, match=msg):
                crb.CudaSurface(size, crb.CudaSurface.FORMAT_RGBA8, n)
        c1 = crb.CudaSurface((64, 64), crb.CudaSurface.FORMAT_RGBA8, 1)
        d1 = crb.CudaSurface((64, 64), crb.CudaSurface.FORMAT_DEPTH32, 1)
        d2 = crb.CudaSurface((64, 32), crb.CudaSurface.FORMAT_DEPTH32, 1)
        d4 = crb.CudaSurface((64, 64), crb.CudaSurface.FORMAT_DEPTH32, 4)
        for args, msg in (((None, d1), "No color buffer"), ((c1, None), "No depth buffer"), ((d1, d1), "Unsupported color buffer format"),
                          ((c1, c1), "Unsupported depth buffer format"), ((c1, d2), "Mismatch in size"), ((c1, d4), "Mismatch in multisampling between surfaces")):
            with pytest.raises(crb.CrbError, match=msg):
                r.setSurfaces(*args)
        with pytest.raises(crb.CrbError, match="Surfaces not set"):
            r.drawTriangles()
        r.setSurfaces(c1, d1)
        with pytest.raises(crb.CrbError, match="Pixel pipe not set"):
            r.drawTriangles()
        with pytest.raises(crb.CrbError, match="Invalid pixel pipe"):
            r.setPixelPipe(None, "PixelPipe_doesNotExist")
        r.setPixelPipe(None, crb.pipe_name("gouraud", 2, 3))
        with pytest.raises(crb.CrbError, match="Vertex buffer not set"):
            r.drawTriangles()
        v, i = crb.scenes.random_soup(50, seed=1, stride_floats=8)
        vb, ib = torch.from_numpy(v).cuda(), torch.from_numpy(i).cuda()
        r.setVertexBuffer(vb, 0)
        with pytest.raises(crb.CrbError, match="Index buffer not set"):
            r.drawTriangles()
        r.setIndexBuffer(ib, 0, i.shape[0])
        with pytest.raises(crb.CrbError, match="Mismatch in multisampling between pixel pipe and surface"):
            r.drawTriangles()
        # the largest surface the format allows: 2048 x 2048, 8 samples; a full-screen triangle, degenerate and repeated-index triangles
        w = h = 2048
        v = np.array([[-1.5, -1.5, 0.3, 1, 1, 0, 0, 1], [1.5, -1.5, 0.3, 1, 0, 1, 0, 1], [0, 1.5, 0.3, 1, 0, 0, 1, 1], [0.2, 0.2, 0.1, 1, 1, 1, 1, 1], [0.21, 0.2, 0.1, 1, 1, 1, 1, 1],
                      [0.2, 0.21, 0.1, 1, 1, 1, 1, 1]], np.float32)
        i = np.array([[0, 1, 2], [3, 3, 4], [3, 4, 5], [5, 4, 3], [0, 0, 0]], np.int32)
        for s_log2, mode in ((3, 0), (0, 2), (0, 0)):
            r.setBinningMode(mode)
            cc, cd = util.draw_cuda(r, crb, v, i, w, h, "gouraud", 3, s_log2)
            g = util.draw_gold(v, i, w, h, "gouraud", 3, s_log2)
            _check_surfaces(cc, cd, g, lsb=1)
            assert (cd != cd[0, 0]).any()
    finally:
        r.close()


def test_tile_major_colour_layout(raster, crb):
    """crb_set_color_layout: the same frame with a tile-major colour surface (what a frame slot in a peer GPU's memory
    uses), on the general, direct and micro paths; depth stays row-major."""
    w, h = 328, 200
    v, i = crb.scenes.grid_gouraud(120, 80)
    try:
        for mode in (0, 2, 3):
            raster.setBinningMode(mode)
            raster.setColorLayout(False)
            cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3)
            raster.setColorLayout(True)
            tc, td = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3)
            rh, rw = cc.shape
            detiled = tc.reshape(rh // 8, rw // 8, 8, 8).transpose(0, 2, 1, 3).reshape(rh, rw)
            assert np.array_equal(detiled, cc) and np.array_equal(td, cd)
        raster.setColorLayout(True)
        with pytest.raises(crb.CrbError, match="single-sample only"):
            util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3, 2)
    finally:
        raster.setColorLayout(False)
        raster.setBinningMode(1)


def test_sort_first_windows_rendered_in_place(raster, crb):
    """crb_set_color_pitch: the four sort-first windows of a frame rendered STRAIGHT into one full-frame image (the
    composite without a paste; across GPUs the image is a PeerFrameSink slot) equal the frame rendered whole."""
    import torch
    fw, fh = 512, 384
    v, i = crb.scenes.random_soup(6000, seed=31, stride_floats=8, size=0.5)
    full_c, full_d = util.draw_cuda(raster, crb, v, i, fw, fh, "gouraud", 3)
    vb, ib = torch.from_numpy(v).cuda(), torch.from_numpy(i).cuda()
    try:
      for mode in (0, 2):   # ordered path, direct path (+ micro-triangles)
        raster.setBinningMode(mode)
        frame = torch.zeros((fh, fw), dtype=torch.int32, device="cuda")
        for (x0, y0, w, h) in [(0, 0, 256, 192), (256, 0, 256, 192), (0, 192, 256, 192), (256, 192, 256, 192)]:
            color = crb.CudaSurface.from_pointer(frame.data_ptr() + 4 * (y0 * fw + x0), (w, h), crb.CudaSurface.FORMAT_RGBA8)
            depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32)
            raster.setSurfaces(color, depth)
            raster.setColorPitch(fw)
            raster.setPixelPipe(None, crb.pipe_name("gouraud", 0, 3))
            raster.setVertexBuffer(vb, 0)
            raster.setIndexBuffer(ib, 0, i.shape[0])
            raster.setSubViewport(fw, fh, x0, y0)
            raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
            raster.drawTriangles()
            assert raster.lastFrameDirect() == (mode == 2)
            assert np.array_equal(depth.numpy(), full_d[y0:y0 + h, x0:x0 + w])
        torch.cuda.synchronize()
        assert np.array_equal(frame.cpu().numpy().view(np.uint32), full_c)
    finally:
        raster.setBinningMode(1)
        raster.setColorPitch(0)
        raster.setSubViewport(0, 0, 0, 0)


def test_async_overflow_is_reported_and_recovered(crb):
    """A fresh context whose first frames are asynchronous and overflow the tile queue (large triangles, tiny initial
    capacity): finish() reports it, the synchronous redraw grows the buffers and the frame is right -- on the ordered path
    and on the direct path (whose per-tile counters and visibility buffer must come back clean)."""
    import torch
    w, h = 1024, 768
    v, i = crb.scenes.random_soup(3000, seed=9, stride_floats=8, size=1.5, clip_fraction=0.0, behind_fraction=0.0)
    g = util.draw_gold(v, i, w, h, "gouraud", 3)
    for mode in (0, 2):
        r = crb.CudaRaster(0)
        try:
            r.setBinningMode(mode)
            color = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8)
            depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32)
            vb, ib = torch.from_numpy(v).cuda(), torch.from_numpy(i).cuda()
            r.setSurfaces(color, depth)
            r.setPixelPipe(None, crb.pipe_name("gouraud", 0, 3))
            r.setVertexBuffer(vb, 0)
            r.setIndexBuffer(ib, 0, i.shape[0])
            for _ in range(3):
                r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
                r.drawTriangles(asynchronous=True)
            with pytest.raises(crb.CrbError, match="overflowed a work buffer"):
                r.finish()
            assert r.getCounters()["overflow"] != 0
            cc, cd = util.draw_cuda(r, crb, v, i, w, h, "gouraud", 3)
            assert r.getCounters()["overflow"] == 0 and r.lastFrameDirect() == (mode == 2)
            _check_surfaces(cc, cd, g, lsb=1)
            for _ in range(2):      # and asynchronous frames are fine from now on
                r.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
                r.drawTriangles(asynchronous=True)
            r.finish()
            _check_surfaces(r._keep["color"].numpy(), r._keep["depth"].numpy(), g, lsb=1)
        finally:
            r.close()


def test_profiling_mode_timers(raster, crb):
    """CR_PROFILING_MODE = ProfilingMode_Timers (reference: cuda/PrivateDefs.hpp:207-270, CudaRaster.cpp:452-487): same
    frame as the plain pipe and a report of per-region percentages that are sane."""
    import re
    w, h = 320, 200
    v, i = crb.scenes.random_soup(8000, seed=78, stride_floats=8, size=0.3)
    cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraudTimers", 3)
    info = raster.getProfilingInfo()
    _check_surfaces(cc, cd, util.draw_gold(v, i, w, h, "gouraud", 3), lsb=1)
    assert "ProfilingMode_Timers" in info and "TriangleSetup:" in info and "FineRaster:" in info
    assert "BinRaster:" in info and "CoarseRaster:" in info
    vals = [float(x) for x in re.findall(r"(-?[0-9.]+)%", info)]
    assert len(vals) == 17 and all(-0.5 <= x <= 100.5 for x in vals), vals
    setup, binr, coarse, fine = vals[:5], vals[5:9], vals[9:12], vals[12:]
    assert sum(setup) > 20.0 and sum(fine) > 20.0 and sum(setup) <= 100.5 and sum(fine) <= 100.5
    assert 20.0 < sum(binr) <= 100.5 and 20.0 < sum(coarse) <= 100.5


def test_sort_first_chunk_bounds_cull(raster, crb):
    """crb_set_chunk_bounds: a sort-first window skips the chunks of 256 triangles whose clip-space box lies outside it.  Frames,
    triSubtris and the surviving records must equal the render without bounds (and the oracle's), on the ordered and the direct
    path; most chunks of a screen-ordered mesh are skipped, and a soup (every chunk spans the screen) loses none."""
    import torch
    fw, fh = 1024, 768
    for name, (v, i) in (("grid", crb.scenes.grid_gouraud(300, 200)), ("soup", crb.scenes.random_soup(20000, seed=3, stride_floats=8, size=0.1))):
        vb, ib = torch.from_numpy(v).cuda(), torch.from_numpy(i).cuda()
        bounds = raster.computeChunkBounds(vb, ib, i.shape[0], 32)
        b = bounds.cpu().numpy()
        assert b.shape == ((i.shape[0] + 255) // 256, 4)
        if name == "grid":
            assert np.isfinite(b).all() and ((b[:, 3] - b[:, 1]) < 0.2).mean() > 0.9      # chunks are thin horizontal strips of the mesh
        try:
            for mode in (0, 3, 1):
                raster.setBinningMode(mode)
                for (x0, y0, w, h) in [(0, 0, 512, 384), (512, 384, 512, 384), (256, 192, 512, 384)]:
                    out = {}
                    for use in (False, True):
                        color = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_RGBA8)
                        depth = crb.CudaSurface((w, h), crb.CudaSurface.FORMAT_DEPTH32)
                        raster.setSurfaces(color, depth)
                        raster.setPixelPipe(None, crb.pipe_name("gouraud", 0, 3))
                        raster.setVertexBuffer(vb, 0)
                        raster.setIndexBuffer(ib, 0, i.shape[0])
                        raster.setChunkBounds(bounds if use else None)
                        raster.setSubViewport(fw, fh, x0, y0)
                        raster.deferredClear((0.2, 0.4, 0.8, 1.0), 1.0)
                        raster.drawTriangles()
                        out[use] = (color.numpy(), depth.numpy(), raster.getWorkBuffers(i.shape[0])["triSubtris"])
                    assert np.array_equal(out[True][0], out[False][0]) and np.array_equal(out[True][1], out[False][1])
                    assert np.array_equal(out[True][2], out[False][2])
                    g = util.draw_gold(v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0))
                    _check_surfaces(out[True][0], out[True][1], g, lsb=1)
                    gs = util.gold_setup(v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0))
                    assert np.array_equal(out[True][2], gs["triSubtris"])
        finally:
            raster.setChunkBounds(None)
            raster.setSubViewport(0, 0, 0, 0)
            raster.setBinningMode(1)
