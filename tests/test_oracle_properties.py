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
