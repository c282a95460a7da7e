"""CPU: size-independent properties of the oracle that the GPU tests rely on (sort-first windows,
idempotence, thread-count independence)."""
import numpy as np

from tests import util


def _windows(fw, fh, cols, rows):
    cw, ch = fw // cols, fh // rows
    return [(c * cw, r * ch, cw, ch) for r in range(rows) for c in range(cols)]


def test_sort_first_windows_equal_the_full_frame(crb):
    """Frame <= 2048 px: every window of any split reproduces the unsplit frame bit for bit,
    including triangles that cross the frustum planes and are clipped."""
    fw, fh = 512, 384
    v, i = crb.scenes.random_soup(6000, seed=31, stride_floats=8, size=0.5)
    full = util.draw_gold(v, i, fw, fh, "gouraud", 3)
    for cols, rows in ((2, 2), (4, 1), (1, 3)):
        for (x0, y0, w, h) in _windows(fw, fh, cols, rows):
            g = util.draw_gold(v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0))
            assert np.array_equal(g["depth"], full["depth"][y0:y0 + h, x0:x0 + w])
            assert np.array_equal(g["color"], full["color"][y0:y0 + h, x0:x0 + w])


def test_4k_frame_is_split_independent(crb):
    """3840x2160 exceeds the 2048 px viewport limit: it is rendered through parent viewports of
    1920x1080; any finer split must give the same pixels."""
    fw, fh = 3840, 2160
    v, i = crb.scenes.random_soup(3000, seed=77, stride_floats=8, size=0.6)
    frames = []
    for cols, rows in ((2, 2), (4, 2), (4, 4)):
        color = np.zeros((fh, fw), np.uint32)
        depth = np.zeros((fh, fw), np.uint32)
        for (x0, y0, w, h) in _windows(fw, fh, cols, rows):
            g = util.draw_gold(v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0))
            color[y0:y0 + h, x0:x0 + w] = g["color"][:h, :w]
            depth[y0:y0 + h, x0:x0 + w] = g["depth"][:h, :w]
        frames.append((color, depth))
    for c, d in frames[1:]:
        assert np.array_equal(c, frames[0][0]) and np.array_equal(d, frames[0][1])
    assert (frames[0][1] != frames[0][1][0, 0]).any()


def test_threads_and_repeat_do_not_change_the_frame(crb):
    w, h = 320, 200
    v, i = crb.scenes.random_soup(4000, seed=3, stride_floats=8, size=0.4)
    a = util.draw_gold(v, i, w, h, "gouraud", 3, 2, "BlendSrcOver", threads=1)
    b = util.draw_gold(v, i, w, h, "gouraud", 3, 2, "BlendSrcOver", threads=7)
    assert np.array_equal(a["color"], b["color"]) and np.array_equal(a["depth"], b["depth"])


def _fullscreen_quad(color_of_x):
    """Two triangles over the whole viewport, w = 1; colour.r = color_of_x(ndc x)."""
    xy = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], np.float32)
    v = np.zeros((4, 8), np.float32)
    v[:, 0:2] = xy
    v[:, 2], v[:, 3] = 0.0, 1.0
    v[:, 4] = [color_of_x(x) for x in xy[:, 0]]
    v[:, 5], v[:, 6], v[:, 7] = 0.0, 0.25, 1.0
    return v, np.array([[0, 1, 2], [0, 2, 3]], np.int32)


def test_quads_mode_derivatives(crb):
    """RenderModeFlag_EnableQuads in the oracle (cuda/PixelPipe.hpp:59-69): dFdx of a colour ramp that grows
    by 1/8 over the 64-pixel-wide frame is 1/512 per pixel everywhere, dFdy is 0, a flat colour gives 0, and
    a shader that takes no derivatives is unaffected by the flag."""
    w, h = 64, 48
    v, i = _fullscreen_quad(lambda x: 0.25 + 0.0625 * x)
    g = util.draw_gold(v, i, w, h, "gouraudQuads", 7)
    r = (g["color"][:h, :w] & 0xFF).astype(np.int64)
    # 8 * |dFdx| * 255 = 8 * (0.125 / 64) * 255 = 3.98 -> 4 (rounded); interior pixels only need the ramp to be affine
    assert set(np.unique(r).tolist()) <= {3, 4, 5} and (r == 4).mean() > 0.9
    assert (((g["color"][:h, :w] >> 8) & 0xFF) == 0).all()          # green ramp is flat
    assert (((g["color"][:h, :w] >> 16) & 0xFF) == 64).all()        # blue passes through
    v2, _ = _fullscreen_quad(lambda x: 0.5)
    g2 = util.draw_gold(v2, i, w, h, "gouraudQuads", 7)
    assert ((g2["color"][:h, :w] & 0xFFFF) == 0).all()
    for s in (0, 2):
        vs, js = crb.scenes.random_soup(3000, seed=12, stride_floats=8, size=0.4)
        a = util.draw_gold(vs, js, 160, 120, "gouraudDiscard", 3, s)
        b = util.draw_gold(vs, js, 160, 120, "gouraudDiscard", 7, s)
        if s == 0:
            assert np.array_equal(a["color"], b["color"]) and np.array_equal(a["depth"], b["depth"])
        else:   # MSAA quads mode also applies the reference's conservative per-pixel kill; depth must still agree
            assert np.array_equal(a["depth"], b["depth"])


def test_quads_mode_msaa_threads_independent(crb):
    vs, js = crb.scenes.random_soup(3000, seed=13, stride_floats=8, size=0.3)
    a = util.draw_gold(vs, js, 160, 120, "gouraudQuads", 7, 2, threads=1)
    b = util.draw_gold(vs, js, 160, 120, "gouraudQuads", 7, 2, threads=5)
    assert np.array_equal(a["color"], b["color"]) and np.array_equal(a["depth"], b["depth"])
    assert len(np.unique(a["color"] & 0xFFFF)) > 20
