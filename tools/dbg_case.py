#!/usr/bin/env python
"""Developer tool (GPU): renders one golden-frame case on every binning mode and reports how many texels differ from the oracle.
    CRB200_LIBRARY=... python tools/dbg_case.py <case index> """
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import cudaraster_linux_b200 as crb
from oracle import gen_golden_frames as gen
from tests import util
case = gen.CASES[int(sys.argv[1])]
name, spec, w, h, shader, flags, s, blend = case
v, i, g = gen.render_case(case)
r = crb.CudaRaster(0)
for mode in (0, 2, 3):
    r.setBinningMode(mode)
    for rep in range(2):
        cc, cd = util.draw_cuda(r, crb, v, i, w, h, shader, flags, s, blend)
        dd = cd != g["depth"]; dc = cc != g["color"]
        ys, xs = np.nonzero(dd)
        print(os.environ.get("CRB200_LIBRARY", "in-tree"), name, "mode", mode, "rep", rep, "depth diffs", int(dd.sum()), "colour diffs", int(dc.sum()),
              "first", list(zip(xs[:4].tolist(), ys[:4].tolist())), [hex(int(cd[y, x])) + "/" + hex(int(g["depth"][y, x])) for x, y in zip(xs[:3], ys[:3])])
