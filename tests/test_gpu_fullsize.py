"""GPU: the BASELINE.json configurations at FULL size against the oracle (VERDICT r1 "parity at BASELINE scale").

C2  1 M-triangle grid, Gouraud, depth, 1920x1080                       (micro-triangle path, direct path, ordered sort)
C3  5 M triangles in 5 shuffled layers, Phong + procedural texture, 4x MSAA, 2048x2048
                                                                        (cuda/FineRaster.inl:855-1126 semantics)
C4  10 M sub-pixel triangles, 1920x1080                                 (between-samples cull, cuda/TriangleSetup.inl:53-95)
C5  (i) 4 M-triangle grid on a 3840x2160 frame as sort-first windows;   (ii) views of the C2 mesh at 1024x1024

Every frame is compared with a LIVE oracle run on all host threads (depth bit-exact, colour <= 1 LSB; exact for constant-colour
pipes) and with the committed fixtures of tests/golden/frames.json (depth CRC32 always; colour CRC32 unless the frame differs
from the oracle by the tolerated LSB).  Setup output (triSubtris / triHeader / triData) is compared bit for bit as well.
"""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import gen_golden_frames as gen
from tests import util

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "frames.json")))
FULL = {c[0]: c for c in gen.FULL_CASES}


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def _threads():
    return max(os.cpu_count() or 8, 1)


def _check_frame(name, cc, cd, g, lsb, what):
    want = GOLD[name]
    assert np.array_equal(cd, g["depth"]), "%s (%s): depth differs from the oracle at %d texels" % (name, what, int((cd != g["depth"]).sum()))
    assert _crc(cd) == want["depth_crc32"], "%s (%s): depth frame differs from the golden fixture" % (name, what)
    if _crc(cc) != want["color_crc32"]:
        assert lsb > 0, "%s (%s): colour frame differs from the golden fixture" % (name, what)
        assert util.color_max_diff(cc, g["color"]) <= lsb, "%s (%s): colour differs from the oracle by more than %d LSB" % (name, what, lsb)


def _run_case(raster, crb, name, modes, lsb, check_setup=True):
    _, spec, w, h, shader, flags, s, blend = FULL[name]
    v, i = gen.scene(spec)
    g = util.draw_gold(v, i, w, h, shader, flags, s, blend, threads=_threads())
    assert _crc(g["depth"]) == GOLD[name]["depth_crc32"] and _crc(g["color"]) == GOLD[name]["color_crc32"]   # the oracle itself
    seen = {}
    try:
        for mode in modes:
            raster.setBinningMode(mode)
            cc, cd = util.draw_cuda(raster, crb, v, i, w, h, shader, flags, s, blend)
            c = raster.getCounters()
            assert c["overflow"] == 0
            seen[mode] = (raster.lastFrameDirect(), c)
            _check_frame(name, cc, cd, g, lsb, "binning mode %d" % mode)
            if check_setup and mode == modes[-1]:
                wb = raster.getWorkBuffers(i.shape[0])
                from oracle import binding as G
                cfg = G.make_config(w, h, s, flags, util.STRIDE[shader], shader, "BlendReplace")
                gs = G.triangle_setup(cfg, v, i, max_subtris=int(g["numSubtris"]) + 64)   # (the default capacity, 7 per triangle, is 5.6 GB for C4)
                util.compare_setup(wb, gs, i.shape[0], flags)
    finally:
        raster.setBinningMode(1)
    return seen, g


def test_c2_full_size_all_paths(raster, crb):
    seen, _ = _run_case(raster, crb, "c2_full_1m_1080p", (2, 3, 0), lsb=1)
    assert seen[2][0] and seen[3][0] and not seen[0][0]
    assert seen[2][1]["numTileEntries"] < seen[3][1]["numTileEntries"]      # mode 2: the micro path took (nearly) everything
    assert seen[0][1]["numBinEntries"] > 900_000


def test_c3_full_size_msaa_phong(raster, crb):
    """5 M triangles, depth complexity ~5 with early-Z kills and overwrites, 4x MSAA, Phong: direct path and ordered sort."""
    seen, g = _run_case(raster, crb, "c3_full_5m_msaa4_2048", (2, 0), lsb=1)
    assert seen[2][0] and not seen[0][0]
    assert seen[2][1]["numTileEntries"] > 5_000_000 and seen[0][1]["numBinEntries"] > 4_500_000
    assert (g["depth"] < 0xFFFFBB3F).mean() > 0.99


def test_c4_full_size_subpixel(raster, crb):
    """10 M sub-pixel triangles: ~85 % die in the between-samples cull of setup; micro path on (2) and off (3), ordered sort (0)."""
    seen, g = _run_case(raster, crb, "c4_full_10m_subpixel_1080p", (2, 3, 0), lsb=0)
    assert seen[2][0] and seen[3][0] and not seen[0][0]
    n_sub = seen[0][1]["numSubtris"] - 10_000_000
    assert n_sub == 0                                                        # nothing is clipped into several sub-triangles here
    assert seen[2][1]["numTileEntries"] == 0 and 1_200_000 < seen[3][1]["numTileEntries"] < 1_700_000
    assert 0.3 < (g["depth"] < 0xFFFFBB3F).mean() < 0.6


@pytest.mark.parametrize("name", ["c5ii_view7_1024", "c5ii_view29_1024"])
def test_c5ii_views_full_size(raster, crb, name):
    seen, _ = _run_case(raster, crb, name, (2, 0), lsb=1, check_setup=False)
    assert seen[2][0]


def test_c5ii_more_views_against_live_oracle(raster, crb):
    """Four more of the 48 views (rotated / scaled / shifted: parts of the mesh leave the frame and are clipped), automatic
    binning mode as bench.py uses it; compared with a live oracle run only."""
    w = h = 1024
    v0, i = crb.scenes.grid_gouraud(1000, 500)
    views = crb.scenes.view_matrix_variants(48)
    for k in (0, 13, 38, 47):
        v = crb.scenes.apply_view(v0, views[k])
        g = util.draw_gold(v, i, w, h, "gouraud", 3, threads=_threads())
        for rep in range(2):   # automatic mode: first frame of the shape on the general path, second on the direct path
            cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3)
            assert np.array_equal(cd, g["depth"]), "view %d: depth differs at %d texels" % (k, int((cd != g["depth"]).sum()))
            assert util.color_max_diff(cc, g["color"]) <= 1


def test_c5i_4k_sort_first_full_size(raster, crb):
    """BASELINE config 5(i): the 4 M-triangle grid on a 3840x2160 frame, four 1920x1080 windows (one parent cell each) and the
    eight 960x1080 windows of an 8-GPU split: every window equals the oracle's render of that window, and the two splits
    compose the same frame."""
    from cudaraster_linux_b200 import multigpu
    fw, fh = 3840, 2160
    v, i = crb.scenes.grid_gouraud_4k()
    frames = {}
    try:
        for parts in (4, 8):
            depth = np.zeros((fh, fw), np.uint32)
            color = np.zeros((fh, fw), np.uint32)
            for (x0, y0, w, h) in multigpu.split_frame(fw, fh, parts):
                cc, cd = util.draw_cuda(raster, crb, v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0))
                if parts == 4:
                    g = util.draw_gold(v, i, w, h, "gouraud", 3, sub=(fw, fh, x0, y0), threads=_threads())
                    assert np.array_equal(cd, g["depth"]), "window (%d, %d): depth differs at %d texels" % (x0, y0, int((cd != g["depth"]).sum()))
                    assert util.color_max_diff(cc, g["color"]) <= 1
                depth[y0:y0 + h, x0:x0 + w] = cd[:h, :w]
                color[y0:y0 + h, x0:x0 + w] = cc[:h, :w]
            frames[parts] = (color, depth)
    finally:
        raster.setSubViewport(0, 0, 0, 0)
    assert np.array_equal(frames[4][0], frames[8][0]) and np.array_equal(frames[4][1], frames[8][1])
    assert (frames[4][1] < 0xFFFFBB3F).mean() > 0.95
