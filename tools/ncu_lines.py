#!/usr/bin/env python
"""Per-source-line executed warp instructions and stall samples of one kernel of an ncu report (developer tool).
    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> [top N]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = "?"
ins = collections.Counter(); smp = collections.Counter(); src = {}
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Name":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0] != "":
        cur = (fname, int(r[0])); src[cur] = r[1].strip()[:110]
    if r[2] == "":
        continue      # source-only row (aggregates): skip, the sass rows below carry the numbers
    try:
        ins[cur] += int(float(r[iE] or 0)); smp[cur] += int(float(r[iS] or 0))
    except ValueError:
        pass
T = sum(ins.values()); S = sum(smp.values())
print("total warp instr", T, "samples", S)
for k, v in ins.most_common(top):
    print("%5.1f%% ins %5.1f%% smp  %s:%d  %s" % (100.0 * v / T, 100.0 * smp[k] / max(S, 1), k[0], k[1], src[k]))
