#!/usr/bin/env python
"""Top stall-sample SASS instructions of a kernel: python scripts/ncu_hotspots.py rep kernel_regex [N]"""
import csv, subprocess, sys, re
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
lines = raw.splitlines()
# first kernel only
idx = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
block = lines[idx[0] + 1: idx[1] if len(idx) > 1 else None]
rows = list(csv.reader(block))
hdr = rows[0]
c = {h: i for i, h in enumerate(hdr)}
data = rows[1:]
tot = sum(int(r[c["# Samples"]]) for r in data)
print("total samples", tot, "instructions", len(data))
# opcode histogram by executed count and samples
from collections import defaultdict
ex = defaultdict(int); sm = defaultdict(int)
for r in data:
    op = r[c["Source"]].split()[0] if not r[c["Source"]].strip().startswith("@") else r[c["Source"]].split()[1]
    op = op.split(".")[0]
    ex[op] += int(r[c["Instructions Executed"]]); sm[op] += int(r[c["# Samples"]])
te = sum(ex.values())
print("opcode: executed share / sample share")
for op, v in sorted(ex.items(), key=lambda kv: -kv[1])[:18]:
    print(f"  {op:10s} {v/te:6.3f} {sm[op]/tot:6.3f}")
stallcols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(int(r[c[h]]) for r in data) for h in stallcols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
print("top instructions by samples:")
for i, r in sorted(enumerate(data), key=lambda ir: -int(ir[1][c["# Samples"]]))[:N]:
    st = sorted(((int(r[c[h]]), h) for h in stallcols), reverse=True)[:2]
    print(f"  #{i:5d} {int(r[c['# Samples']]):6d} {r[c['Source']].strip()[:70]:70s} {st}")
