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
    ppm = open(out + ".ppm", "rb").read()       # CudaSurface::resolveToFile (stand-in for resolveToScreen)
    assert ppm.startswith(b"P6\n%d %d\n255\n" % (w, h)) and len(ppm) == len(b"P6\n%d %d\n255\n" % (w, h)) + w * h * 3
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


def test_cpp_runtime_compiled_pipe(tmp_path):
    """FW::CudaCompiler (the reference's run-time shader compilation, gpu/CudaCompiler.cpp): the user pipe
    source is compiled by nvcc at run time with a -D define that changes the shader, cached on disk, and the
    second run is a cache hit; the frame is checked against the oracle."""
    exe = os.path.join(ROOT, "examples", "cpp", "cube")
    src = os.path.join(ROOT, "examples", "cpp", "UserPipes.cu")
    out = str(tmp_path / "cube.raw")
    w, h, shift = 640, 480, 4
    r = subprocess.run([exe, src, out, str(w), str(h), str(shift)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "CudaCompiler: Compiling" in r.stdout and "Done." in r.stdout
    r2 = subprocess.run([exe, src, out, str(w), str(h), str(shift)], capture_output=True, text=True, timeout=120)
    assert r2.returncode == 0 and "cached" in r2.stdout, r2.stdout + r2.stderr
    raw = np.fromfile(out, np.uint32)
    tw, th = int(raw[0]), int(raw[1])
    color = raw[2:2 + tw * th].reshape(th, tw)
    depth = raw[2 + tw * th:2 + 2 * tw * th].reshape(th, tw)
    verts = raw[2 + 2 * tw * th:].view(np.float32).reshape(8, 8)
    idx = np.array([[7, 3, 1], [7, 1, 5], [7, 5, 6], [6, 5, 4], [6, 4, 2], [2, 4, 0], [2, 0, 3], [3, 0, 1], [3, 7, 6], [3, 6, 2], [5, 1, 0], [5, 0, 4]], np.int32)
    g = util.draw_gold(verts, idx, w, h, "gouraud", 3)
    assert np.array_equal(depth, g["depth"])
    covered = depth < 0xFFFFBB3F
    yy, xx = np.mgrid[0:th, 0:tw]
    plain = covered & ((((xx >> shift) ^ (yy >> shift)) & 1) == 0)      # 16-pixel checker: the define reached the shader
    assert util.color_max_diff(color[plain], g["color"][plain]) <= 1
    dark = covered & ~plain
    assert dark.sum() > 1000 and ((color[dark] & 0xFF) <= (g["color"][dark] & 0xFF)).all()
    # a bad source fails like the reference: message + non-zero exit
    bad = tmp_path / "Bad.cu"
    bad.write_text("#include <cudaraster/cuda/PixelPipe.inl>\nthis is not CUDA\n")
    r3 = subprocess.run([exe, str(bad), out, str(w), str(h)], capture_output=True, text=True, timeout=600)
    assert r3.returncode != 0 and "CudaCompiler: Compilation of" in (r3.stdout + r3.stderr)


@pytest.mark.parametrize("world", [2, 4])
def test_cpp_sort_first_multi_process(world):
    """include/cudaraster/MultiGpu.hpp: `world` PROCESSES (one per GPU; they share GPUs when the box has fewer) render one
    2560x1440 frame sort-first, in place in rank 0's memory through CUDA IPC; rank 0 compares with its own render of the frame."""
    exe = os.path.join(ROOT, "examples", "cpp", "sortfirst")
    assert os.path.exists(exe), "examples/cpp is not built (run __graft_entry__.build())"
    r = subprocess.run([exe, str(world)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "-> OK" in r.stdout, r.stdout + r.stderr
    assert r.stdout.count("rendered") == world
