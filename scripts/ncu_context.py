#!/usr/bin/env python
"""SASS context around given instruction indices, with the CUDA source line each instruction belongs to:
python scripts/ncu_context.py rep kernel_regex idx[,idx...] [radius]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
want = [int(x) for x in sys.argv[3].split(",")]
rad = int(sys.argv[4]) if len(sys.argv) > 4 else 6
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# build: address -> (file, line, src)
addr2src = {}
fname = ""; hdr = None; cur = None; nfunc = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": nfunc += 1; continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None: continue
    if r[0] != "": cur = (fname, r[0], r[1].strip()[:100]); continue
    if len(r) > 3 and r[2].startswith("0x") and cur and r[2] not in addr2src: addr2src[r[2]] = cur
raw2 = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
lines = raw2.splitlines()
idx = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
block = lines[idx[0] + 1: idx[1] if len(idx) > 1 else None]
data = list(csv.reader(block)); h = {k: i for i, k in enumerate(data[0])}; data = data[1:]
for w in want:
    print("=" * 100)
    for i in range(max(0, w - rad), min(len(data), w + rad + 1)):
        r = data[i]
        src = addr2src.get(r[h["Address"]], ("", "", ""))
        print(f"{'>>' if i == w else '  '} #{i:5d} {int(r[h['# Samples']]):6d} {r[h['Source']].strip()[:60]:60s} | {src[0]}:{src[1]} {src[2][:70]}")
