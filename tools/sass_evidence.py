#!/usr/bin/env python
"""Counts, per kernel of libcrb200.so, the SASS mnemonics that show which sm_100 mechanisms the hand-written kernels use
(developer tool; writes profiles/<tag>_sass_evidence.md):  python tools/sass_evidence.py r2"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = os.path.join(ROOT, "cudaraster-linux_b200", "libcrb200.so")
WHAT = [("UBLKCP", "cp.async.bulk shared -> global (bulk-copy engine, TMA path)"),
        ("ENL2.256.CONSTANT", "256-bit global load (ld.global.nc[.L1::no_allocate].v8.b32 -> LDG.E[.NA].ENL2.256.CONSTANT)"),
        ("STG.E.ENL2.256", "256-bit global store (st.global.v8.b32)"),
        ("REDG.E.MIN.64", "red.global.min.u64 (visibility-buffer write of the micro raster)"),
        ("UCGABAR_ARV", "cluster barrier arrive (cluster.sync of the DSMEM bin scan)"),
        ("UCGABAR_WAIT", "cluster barrier wait"),
        ("ACQBULK", "griddepcontrol.wait (programmatic dependent launch)"),
        ("PREEXIT", "griddepcontrol.launch_dependents"),
        ("REDUX", "redux.sync warp reductions"),
        ("ATOMG", "global atomics with a result"),
        ("REDG.E.ADD", "fire-and-forget global reductions (tile counters)")]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn = None
cnt = collections.defaultdict(collections.Counter)
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = re.sub(r"\(anonymous namespace\)::|FW::|void ", "", fn)
        fn = re.sub(r"\(crb_frame\)$", "", fn)
        continue
    for k, _ in WHAT:
        if k in ln:
            cnt[k][fn] += 1
lines = ["# %s -- SASS evidence of the sm_100 mechanisms in libcrb200.so (cuobjdump -sass, nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)" % tag, "",
         "Made by `tools/sass_evidence.py`; counts are static instruction counts per kernel instance.", ""]
for k, what in WHAT:
    tot = sum(cnt[k].values())
    lines.append("## `%s` -- %s: %d instructions in %d kernels" % (k, what, tot, len(cnt[k])))
    for f, n in sorted(cnt[k].items(), key=lambda kv: (-kv[1], kv[0]))[:6]:
        lines.append("* %d  `%s`" % (n, f[:150]))
    if len(cnt[k]) > 6:
        lines.append("* ... %d more" % (len(cnt[k]) - 6))
    lines.append("")
path = os.path.join(ROOT, "profiles", "%s_sass_evidence.md" % tag)
open(path, "w").write("\n".join(lines))
print("wrote", path)
