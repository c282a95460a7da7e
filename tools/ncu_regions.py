#!/usr/bin/env python
"""Rolls the per-line output of tools/ncu_lines.py up into named source regions (developer tool).
    python tools/ncu_regions.py <report> <kernel> <lib> <file> name:lo-hi name:lo-hi ..."""
import collections, subprocess, sys, os
rep, kern, lib, fname = sys.argv[1:5]
groups = []
for g in sys.argv[5:]:
    name, rng = g.split(":")
    lo, hi = rng.split("-")
    groups.append((name, int(lo), int(hi)))
out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "ncu_lines.py"), rep, kern, "--lib", lib, "--top", "100000"], capture_output=True, text=True).stdout
lines = out.splitlines()
print("\n".join(lines[:3]))
acc = collections.OrderedDict((g[0], [0, 0]) for g in groups)
other = collections.defaultdict(lambda: [0, 0])
tot = ts = 0
for ln in lines[4:]:
    r = ln.split()
    if len(r) < 6 or ":" not in r[0]:
        continue
    f, l = r[0].rsplit(":", 1)
    n, sm = int(r[1]), int(r[3])
    tot += n; ts += sm
    for name, lo, hi in groups:
        if f == fname and lo <= int(l) <= hi:
            acc[name][0] += n; acc[name][1] += sm
            break
    else:
        other[f][0] += n; other[f][1] += sm
for k, v in list(acc.items()) + sorted(other.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %10d %5.1f%%  samples %5.1f%%" % (k, v[0], 100 * v[0] / max(tot, 1), 100 * v[1] / max(ts, 1)))
