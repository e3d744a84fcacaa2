#!/usr/bin/env python
"""Turns gpurun_out/{launches.csv,prof.ncu-rep} into tracked summaries under profiles/ (prof.ncu-rep holds one
launch of beam_forward and one of beam_backward: the two launches of a Newton iteration).
    python scripts/make_profile_summary.py r01
"""
import csv, json, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
# ---- launch list
rows = [r for r in csv.reader(l for l in open(os.path.join(ROOT, "gpurun_out", "launches.csv")) if l.startswith('"'))]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
d = defaultdict(list)
for r in rows[1:]:
    d[r[ki]].append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in d.values())
with open(os.path.join(out, f"{tag}_launch_list.md"), "w") as f:
    f.write(f"# {tag}: launch list (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n")
    f.write("Command: `python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline` (65,536 beams x 200 cells, first 60 launches).\n")
    f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|\n")
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k[:70]}` | {len(v)} | {sum(v)/1e6:.3f} | {sum(v)/len(v)/1e6:.4f} | {sum(v)/tot:.4f} |\n")
# ---- full capture
raw = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out", "prof.ncu-rep"), "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]; col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
summary = []
with open(os.path.join(out, f"{tag}_newton_iteration_ncu.md"), "w") as f:
    f.write(f"# {tag}: ncu --set full --clock-control none, the two launches of one Newton iteration: beam_forward<BFK_NEWMARK>, beam_backward<BFK_NEWMARK> (shown by ncu as <1>)\n\n")
    for r in rows[2:]:
        f.write(f"## launch: {r[col['Kernel Name']]}  grid {r[col.get('Grid Size', 0)]} block {r[col.get('Block Size', 0)]}\n\n| metric | value | unit |\n|---|---|---|\n")
        rec = {}
        for w in want:
            if w in col:
                f.write(f"| {w} | {r[col[w]]} | {rows[1][col[w]]} |\n"); rec[w] = r[col[w]]
        st = []
        for h, i in col.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try: st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError: pass
        f.write("\nWarp stalls per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]) + "\n\n")
        summary.append(rec)
def gb(x, unit):
    x = float(x); return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(unit, 1)
unit = lambda k: rows[1][col[k]]
per = {}
for r, row in zip(summary, rows[2:]):
    name = "beam_forward" if "beam_forward" in row[col["Kernel Name"]] else "beam_backward" if "beam_backward" in row[col["Kernel Name"]] else row[col["Kernel Name"]]
    per[name] = gb(r["dram__bytes_read.sum"], unit("dram__bytes_read.sum")) + gb(r["dram__bytes_write.sum"], unit("dram__bytes_write.sum"))
traffic = per.get("beam_forward", 0) + per.get("beam_backward", 0)
# ---- FP64 work per launch from the source page: thread-level DFMA (2 flop), DMUL, DADD (1 flop)
def fp64_flop(kernel):
    raw = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out", "prof.ncu-rep"), "--page", "source", "--csv",
                          "--kernel-name", f"regex:{kernel}"], capture_output=True, text=True).stdout
    lines = raw.splitlines()
    idx = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
    if not idx:
        return None
    block = list(csv.reader(lines[idx[0] + 1: idx[1] if len(idx) > 1 else None]))
    c = {h: i for i, h in enumerate(block[0])}
    flop = 0
    for r in block[1:]:
        src = r[c["Source"]].split()
        if not src:
            continue
        op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
        w = {"DFMA": 2, "DMUL": 1, "DADD": 1}.get(op)
        if w:
            flop += w * int(r[c["Thread Instructions Executed"]])
    return flop
flops = {k: fp64_flop(k) for k in ("beam_forward", "beam_backward")}
json.dump({"newton_iteration_dram_bytes": traffic, "beam_forward_dram_bytes": per.get("beam_forward"), "beam_backward_dram_bytes": per.get("beam_backward"),
           "beam_forward_fp64_flop": flops["beam_forward"], "beam_backward_fp64_flop": flops["beam_backward"],
           "source": f"profiles/{tag}_newton_iteration_ncu.md", "cells_per_launch": 13107200, "bytes_per_update": traffic / 13107200,
           "fp64_flop_per_update": (flops["beam_forward"] + flops["beam_backward"]) / 13107200 if all(flops.values()) else None},
          open(os.path.join(out, "traffic.json"), "w"), indent=1)
print("traffic per Newton iteration %.3f GB = %.0f B/update" % (traffic / 1e9, traffic / 13107200), per, "fp64 flop", flops)
