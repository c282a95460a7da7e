#!/usr/bin/env python
"""Per-source-line summary of one kernel of an ncu report (developer tool, not product).

    python tools/ncu_lines.py gpurun_out/c2_full.ncu-rep fineRasterSingleKernel [--top 40]

ncu's CSV source page is per SASS instruction and carries no line numbers; this joins it (by
instruction order) with `nvdisasm -g` of the same kernel taken from the in-tree library, and
prints executed warp instructions and stall samples per CUDA source line (innermost inlined
location), plus a per-file/function roll-up.  The library must be the build the report was taken from.
"""
import argparse
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(lib, kernel_re):
    tmp = tempfile.mkdtemp()
    lib = os.path.abspath(lib)
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, check=True, capture_output=True)
    out = []
    for cubin in sorted(glob.glob(os.path.join(tmp, "*.cubin"))):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        cur_fn, cur_loc, stack = None, None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", ln)
            if m:
                cur_fn = m.group(1)
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
            if m:
                cur_loc = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and cur_fn and re.search(kernel_re, cur_fn):
                out.append((cur_fn, int(m.group(1), 16), m.group(2).strip(), cur_loc))
        if out:
            break
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--lib", default=os.path.join(ROOT, "cudaraster-linux_b200", "libcrb200.so"))
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--id", type=int, default=None, help="ncu launch id inside the report, if the name matches several")
    args = ap.parse_args()
    cmd = ["ncu", "-i", args.report, "--page", "source", "--csv", "--kernel-name", "regex:" + args.kernel]
    r = subprocess.run(cmd, capture_output=True, text=True)
    rows = list(csv.reader(r.stdout.splitlines()))
    # the page may hold several kernels: take the first block
    hdr_idx = [i for i, row in enumerate(rows) if row and row[0] == "Address"]
    if not hdr_idx:
        sys.exit("no source page: " + r.stdout[:300] + r.stderr[:300])
    blk = 0 if args.id is None else args.id
    start = hdr_idx[blk]
    end = hdr_idx[blk + 1] - 1 if blk + 1 < len(hdr_idx) else len(rows)
    hdr = rows[start]
    body = [row for row in rows[start + 1:end] if len(row) == len(hdr)]
    name = rows[start - 1][1] if start > 0 else args.kernel
    col = {k: i for i, k in enumerate(hdr)}
    kernel_name_re = re.escape(args.kernel)
    sass = sass_lines(args.lib, kernel_name_re)
    # several template instances can match: pick the one whose instruction count equals the report's
    by_fn = collections.OrderedDict()
    for fn, addr, txt, loc in sass:
        by_fn.setdefault(fn, []).append((addr, txt, loc))
    cand = [fn for fn, v in by_fn.items() if len(v) == len(body)]
    if not cand:
        sys.exit("no SASS function with %d instructions (have %s)" % (len(body), {k: len(v) for k, v in by_fn.items()}))
    ins = by_fn[cand[0]]
    print("kernel:", name[:110])
    print("sass fn:", cand[0][:110], "instructions:", len(ins))
    per_line = collections.defaultdict(lambda: [0, 0, 0])
    tot_inst = tot_samp = 0
    for row, (addr, txt, loc) in zip(body, ins):
        ex = int(float(row[col["Instructions Executed"]] or 0))
        sm = int(float(row[col["# Samples"]] or 0)) if "# Samples" in col else 0
        per_line[loc][0] += ex
        per_line[loc][1] += sm
        per_line[loc][2] += 1
        tot_inst += ex
        tot_samp += sm
    print("total warp instructions executed: %d, stall samples: %d" % (tot_inst, tot_samp))
    print("%-28s %12s %6s %9s %6s %5s" % ("file:line", "warp-inst", "%", "samples", "%", "sass"))
    for loc, (ex, sm, n) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:args.top]:
        print("%-28s %12d %6.2f %9d %6.2f %5d" % ("%s:%d" % loc if loc else "?", ex, 100.0 * ex / max(tot_inst, 1), sm, 100.0 * sm / max(tot_samp, 1), n))


if __name__ == "__main__":
    main()
