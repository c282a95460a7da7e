#!/usr/bin/env python
"""Executed-instruction mix of one kernel of an ncu report by SASS opcode and issue pipe (developer tool).
    python tools/ncu_opmix.py <report.ncu-rep> <kernel regex>
fma pipe: IMAD / FFMA / FMUL / FADD ...; alu pipe: IADD3 / LOP3 / SHF / ISETP / SEL ... (each pipe issues one warp
instruction every 2 cycles per scheduler: a kernel whose mix is one-sided is bound at 2 x that share)."""
import collections, csv, re, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)
hdr = rows[hi]; col = {k: i for i, k in enumerate(hdr)}
tot = collections.Counter()
for r in rows[hi + 1:end]:
    if len(r) != len(hdr):
        continue
    try:
        ex = int(float(r[col["Instructions Executed"]] or 0))
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
    tot[(m.group(2) if m else "?").split(".")[0]] += ex
T = sum(tot.values())
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2"}
ALU = {"IADD3", "IADD", "LOP3", "SHF", "PRMT", "FMNMX", "ISETP", "SEL", "LEA", "FSETP", "MOV", "IABS", "VIMNMX", "VIADD", "FSEL", "PLOP3", "IMNMX", "SGXT", "BMSK", "FSET", "VABSDIFF", "VIADDMNMX", "CS2R", "S2R"}
XU = {"MUFU", "F2I", "I2F", "I2FP", "F2F", "POPC", "FLO", "BREV", "F2FP", "F2IP", "FCHK"}
LSU = {"LDG", "STG", "LDS", "STS", "ATOMG", "ATOMS", "RED", "LDL", "STL", "LDC", "ATOM", "LDGSTS", "LDSM"}
g = lambda S: sum(v for k, v in tot.items() if k in S)
print("total warp instructions", T)
for name, S in (("fma pipe", FMA), ("alu pipe", ALU), ("xu", XU), ("lsu", LSU)):
    print("%-9s %10d %5.1f%%" % (name, g(S), 100.0 * g(S) / T))
print("other     %10d %5.1f%%" % (T - g(FMA) - g(ALU) - g(XU) - g(LSU), 100.0 * (T - g(FMA) - g(ALU) - g(XU) - g(LSU)) / T))
for k, v in tot.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    print("  %-10s %10d %5.1f%%" % (k, v, 100.0 * v / T))
