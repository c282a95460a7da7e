#!/usr/bin/env python
"""Turns an `ncu --set full` report of the bench into the tracked summaries under profiles/
(developer tool):

    python tools/summarize_ncu.py gpurun_out/c2_full_f.ncu-rep c2 r1

writes profiles/<round>_ncu_<workload>.md (one row per kernel: duration, DRAM bytes, issue
utilisation, occupancy, cache hit rates, top stall reasons) and updates profiles/traffic.json
({workload: {stage: dram bytes per launch}}), which bench.py copies into roofline.traffic.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGE_OF = {"triangleSetupKernel": "triangleSetup", "binScanKernel": "binRaster", "binScatterKernel": "binRaster", "coarseScanKernel": "coarseRaster",
            "coarseScatterKernel": "coarseRaster", "directAllocKernel": "binRaster", "directScatterKernel": "coarseRaster", "fineRasterSingleKernel": "fineRaster", "fineRasterMultiKernel": "fineRaster"}


def main():
    rep, workload, rnd = sys.argv[1], sys.argv[2], sys.argv[3]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {k: i for i, k in enumerate(hdr)}

    def val(r, k, scale=1.0):
        try:
            return float(r[col[k]]) * scale
        except Exception:
            return float("nan")

    def to_bytes(r, k):
        u = units[col[k]]
        return val(r, k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)

    def to_us(r, k):
        u = units[col[k]]
        return val(r, k) * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)

    stall_keys = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k]
    lines = ["# %s -- ncu --set full, workload %s (one launch per kernel, --clock-control none)" % (rnd, workload), "",
             "Cold-cache, serialised launches: use the SHARES, not the absolute times (bench.py times the kernels live with CUDA events).", "",
             "| kernel | grid x block | regs | time us | DRAM read MB | DRAM write MB | issue slots busy % | warps active % of max | L1 hit % | L2 hit % | top stalls (warps per issue) |",
             "|---|---|---|---|---|---|---|---|---|---|---|"]
    traffic = {}
    for r in body:
        name = r[col["Kernel Name"]]
        short = next((k for k in STAGE_OF if k in name), name.split("(")[0][:40])
        rd, wr = to_bytes(r, "dram__bytes_read.sum"), to_bytes(r, "dram__bytes_write.sum")
        stalls = sorted(((val(r, k), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in stall_keys), reverse=True)[:3]
        lines.append("| %s | %s x %s | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %.1f | %s |" % (
            short, r[col["Grid Size"]], r[col["Block Size"]], int(val(r, "launch__registers_per_thread")), to_us(r, "gpu__time_duration.sum"), rd / 1e6, wr / 1e6,
            val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active"), val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
            val(r, "l1tex__t_sector_hit_rate.pct"), val(r, "lts__t_sector_hit_rate.pct"), ", ".join("%s %.2f" % (b, a) for a, b in stalls)))
        stage = STAGE_OF.get(short)
        if stage:
            traffic[stage] = traffic.get(stage, 0) + int(rd + wr)
    lines += ["", "DRAM bytes per launch, summed per stage (-> profiles/traffic.json -> bench.py roofline.traffic): " + json.dumps(traffic), ""]
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "%s_ncu_%s.md" % (rnd, workload)), "w") as fh:
        fh.write("\n".join(lines))
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    allt = json.load(open(tp)) if os.path.exists(tp) else {}
    allt[workload] = traffic
    json.dump(allt, open(tp, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
