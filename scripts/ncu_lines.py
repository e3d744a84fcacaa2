#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of a kernel: python scripts/ncu_lines.py rep kernel_regex [N]"""
import csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
out = []; fname = ""; hdr = None; seen_kernel = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        seen_kernel += 1
        continue
    if r[0] == "Line No": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or r[0] == "": continue
    try:
        ex = int(r[hdr["Instructions Executed"]]); sm = int(r[hdr["# Samples"]])
    except (ValueError, IndexError):
        continue
    out.append((ex, sm, fname, r[0], r[1].strip()[:110]))
tot = sum(o[0] for o in out); tots = sum(o[1] for o in out)
print("total executed", tot, "samples", tots)
for ex, sm, f, ln, src in sorted(out, reverse=True)[:N]:
    print(f"{ex/tot:6.3f} {sm/max(tots,1):6.3f} {f}:{ln:>4s}  {src}")
