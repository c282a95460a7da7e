"""GPU: the kept C++ host API (include/cudaraster/CudaRaster.hpp) and the device plugin API
(CR_DEFINE_PIXEL_PIPE in a USER shared object, resolved by name) -- examples/cpp renders the
reference demo's cube; the frame is checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_cube_with_user_pipe_module(tmp_path):
    exe = os.path.join(ROOT, "examples", "cpp", "cube")
    mod = os.path.join(ROOT, "examples", "cpp", "libuserpipes.so")
    assert os.path.exists(exe) and os.path.exists(mod), "examples/cpp is not built (run __graft_entry__.build())"
    out = str(tmp_path / "cube.raw")
    w, h = 1024, 768
    r = subprocess.run([exe, mod, out, str(w), str(h)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "triangleSetup" in r.stdout and "fine =" in r.stdout
    raw = np.fromfile(out, np.uint32)
    tw, th = int(raw[0]), int(raw[1])
    assert (tw, th) == (w, h)
    color = raw[2:2 + tw * th].reshape(th, tw)
    depth = raw[2 + tw * th:2 + 2 * tw * th].reshape(th, tw)
    verts = raw[2 + 2 * tw * th:].view(np.float32).reshape(8, 8)
    idx = np.array([[7, 3, 1], [7, 1, 5], [7, 5, 6], [6, 5, 4], [6, 4, 2], [2, 4, 0], [2, 0, 3], [3, 0, 1], [3, 7, 6], [3, 6, 2], [5, 1, 0], [5, 0, 4]], np.int32)
    g = util.draw_gold(verts, idx, w, h, "gouraud", 3)
    assert np.array_equal(depth, g["depth"])
    covered = depth < 0xFFFFBB3F
    assert 0.05 < covered.mean() < 0.7 and (color[~covered] == 0xFFCC6633).all()
    yy, xx = np.mgrid[0:th, 0:tw]
    plain = covered & ((((xx >> 3) ^ (yy >> 3)) & 1) == 0)      # the user shader leaves these pixels as plain Gouraud
    assert util.color_max_diff(color[plain], g["color"][plain]) <= 1
    dark = covered & ~plain
    assert ((color[dark] & 0xFF) <= (g["color"][dark] & 0xFF)).all() and (color[dark] >> 24 == 0xFF).all()
