#!/usr/bin/env python
"""Developer tool (GPU): setup records of one golden-frame case against the oracle, mismatches by word position.
    CRB200_LIBRARY=... python tools/dbg_setup.py <case index>"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np
import cudaraster_linux_b200 as crb
from oracle import gen_golden_frames as gen
from tests import util
case = gen.CASES[int(sys.argv[1])]
name, spec, w, h, shader, flags, s, blend = case
v, i = gen.scene(spec)
gs = util.gold_setup(v, i, w, h, shader, flags, s)
r = crb.CudaRaster(0)
r.setBinningMode(0)
util.draw_cuda(r, crb, v, i, w, h, shader, flags, s, blend)
wb = r.getWorkBuffers(i.shape[0])
cs = wb["triSubtris"]; single = np.nonzero((cs == 1) & (gs["triSubtris"] == 1))[0]
cd, gd = wb["triData"][single], gs["triData"][single]
print(os.environ.get("CRB200_LIBRARY", "in-tree"), name, "singles", len(single), "word mismatches", (cd != gd).sum(axis=0).tolist())
k = np.nonzero((cd != gd).any(axis=1))[0][:3]
for t in k:
    print(" tri", int(single[t]), "cuda", [hex(int(x)) for x in cd[t]], "\n      gold", [hex(int(x)) for x in gd[t]])
