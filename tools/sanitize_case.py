#!/usr/bin/env python
"""Developer tool: a few small frames on every binning path, meant to run under compute-sanitizer
(`compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import cudaraster_linux_b200 as crb  # noqa: E402
from tests import util  # noqa: E402

r = crb.CudaRaster(0)
w, h = 200, 136
small = crb.scenes.grid_gouraud(60, 40)
soup = crb.scenes.random_soup(1500, seed=3, stride_floats=8)
for mode in (0, 2, 3):
    r.setBinningMode(mode)
    for name, (v, i) in (("grid", small), ("soup", soup)):
        for shader, flags, s, blend in (("gouraud", 3, 0, "BlendReplace"), ("gouraud", 3, 2, "BlendReplace"), ("gouraud", 3, 0, "BlendSrcOver"), ("gouraudQuads", 7, 0, "BlendReplace")):
            cc, cd = util.draw_cuda(r, crb, v, i, w, h, shader, flags, s, blend)
            g = util.draw_gold(v, i, w, h, shader, flags, s, blend)
            ok = np.array_equal(cd, g["depth"]) and util.color_max_diff(cc, g["color"]) <= 1
            print(mode, name, shader, s, blend, "direct" if r.lastFrameDirect() else "general", "ok" if ok else "MISMATCH")
r.close()
