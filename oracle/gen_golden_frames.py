#!/usr/bin/env python
"""Generates tests/golden/frames.json: CRC32 checksums of whole frames rendered by the oracle (test infrastructure) on
seeded scenes -- fixtures that pin the oracle against regressions and let the GPU tests compare with committed values
instead of a live oracle run only.  Usage: python oracle/gen_golden_frames.py"""
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import cudaraster_linux_b200 as crb  # noqa: E402
from tests import util  # noqa: E402

CASES = [
    # name, scene, width, height, shader, flags, samplesLog2, blend
    ("c1_cube_1024x768", ("cube", {"width": 1024, "height": 768}), 1024, 768, "passthrough", 1, 0, "BlendReplace"),
    ("c1_cube_720x480", ("cube", {"width": 720, "height": 480}), 720, 480, "passthrough", 1, 0, "BlendReplace"),
    ("grid_gouraud_512x384", ("grid_gouraud", {"nx": 160, "ny": 100}), 512, 384, "gouraud", 3, 0, "BlendReplace"),
    ("soup_gouraud_msaa4_320x200", ("random_soup", {"num_tris": 8000, "seed": 101, "stride_floats": 8}), 320, 200, "gouraud", 3, 2, "BlendReplace"),
    ("soup_srcover_256x192", ("random_soup", {"num_tris": 5000, "seed": 4242, "stride_floats": 8, "size": 0.5}), 256, 192, "gouraud", 3, 0, "BlendSrcOver"),
    ("soup_quads_256x192", ("random_soup", {"num_tris": 4000, "seed": 515, "stride_floats": 8, "size": 0.5}), 256, 192, "gouraudQuads", 7, 0, "BlendReplace"),
]
# The BASELINE.json configurations at FULL size: C2, C3, C4 and a view of C5(ii).  The oracle renders each in
# seconds on all host threads; the GPU tests (tests/test_gpu_fullsize.py) compare with these fixtures AND with a live oracle run.
FULL_CASES = [
    ("c2_full_1m_1080p", ("grid_gouraud", {"nx": 1000, "ny": 500}), 1920, 1080, "gouraud", 3, 0, "BlendReplace"),
    ("c3_full_5m_msaa4_2048", ("layered_phong", {"nx": 1000, "ny": 500, "layers": 5}), 2048, 2048, "texPhong", 3, 2, "BlendReplace"),
    ("c4_full_10m_subpixel_1080p", ("subpixel_soup", {"num_tris": 10_000_000}), 1920, 1080, "passthrough", 1, 0, "BlendReplace"),
    ("c5ii_view7_1024", ("grid_gouraud_view", {"nx": 1000, "ny": 500, "view": 7}), 1024, 1024, "gouraud", 3, 0, "BlendReplace"),
    ("c5ii_view29_1024", ("grid_gouraud_view", {"nx": 1000, "ny": 500, "view": 29}), 1024, 1024, "gouraud", 3, 0, "BlendReplace"),
]


def scene(spec):
    fn, kw = spec
    v, i = getattr(crb.scenes, fn)(**kw)
    if fn == "cube":
        return v, i
    return v, i


def render_case(case, threads=8):
    name, spec, w, h, shader, flags, s, blend = case
    v, i = scene(spec)
    if shader == "passthrough":
        v = np.ascontiguousarray(v[:, :4])
    g = util.draw_gold(v, i, w, h, shader, flags, s, blend, threads=threads)
    return v, i, g


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def main():
    out = {}
    for case in CASES + FULL_CASES:
        v, i, g = render_case(case)
        out[case[0]] = {"depth_crc32": crc(g["depth"]), "color_crc32": crc(g["color"]), "shape": list(g["depth"].shape), "numSubtris": int(g["numSubtris"])}
        print(case[0], out[case[0]])
    path = os.path.join(ROOT, "tests", "golden", "frames.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
