#!/usr/bin/env python
"""Turns gpurun_out/{launches.csv, prof_<family><beams>.ncu-rep} (scripts/gpu_profile_r02.sh) into the tracked summaries under
profiles/ and into profiles/traffic.json (measured DRAM bytes and FP64 work per launch, which bench.py scales to its
bundle; stamped with the hash of the kernel sources so that stale numbers are never reported).
    python scripts/make_profile_summary.py r02
"""
import csv, json, os, subprocess, sys
from collections import defaultdict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_sources_sha  # noqa: E402
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)
CELLS = 200
CAPTURES = {  # report -> (family key of traffic.json, beams, forward kernels, backward kernels)
    "split65536": ("split", 65536, ["beam_forward_split"], ["beam_backward_split"]),
    "perbeam32768": ("per_beam", 32768, ["beam_forward"], ["beam_backward"]),
    "splitcell8192": ("split_cell", 8192, ["beam_forward_split"], ["beam_backsub_split", "beam_update"]),
    "splitgroup4096": ("split_group", 4096, ["beam_forward_ws"], ["beam_backsub_split", "beam_update"]),
    "group4096": ("group", 4096, ["beam_forward_ws"], ["beam_backsub", "beam_update"]),
}
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}

# ---- launch list of the default bench command
path = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = defaultdict(list)
    for r in rows[1:]:
        d[r[ki]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    with open(os.path.join(out, f"{tag}_launch_list.md"), "w") as f:
        f.write(f"# {tag}: launch list (ncu --metrics gpu__time_duration.sum --clock-control none)\n\n")
        f.write("Command: `python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-verify` (65,536 beams x 200 cells on one B200, "
                "first 120 launches, BEAMGPU_NO_PIPELINE=1).\nPer-launch times under ncu are cold-cache and serialised: compare SHARES.\n\n"
                "| kernel | launches | total ms | avg ms | share |\n|---|---|---|---|---|\n")
        for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k[:80]}` | {len(v)} | {sum(v)/1e6:.3f} | {sum(v)/len(v)/1e6:.4f} | {sum(v)/tot:.4f} |\n")

def fp64_flop(rep, kernel):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:^{kernel}$"], capture_output=True, text=True).stdout
    lines = raw.splitlines()
    idx = [i for i, l in enumerate(lines) if l.startswith('"Kernel Name"')]
    if not idx:
        return 0
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

traffic = {"kernel_sources_sha": kernel_sources_sha(), "source": f"profiles/{tag}_*_ncu.md (ncu --set full --clock-control none, one launch per kernel)"}
for name, (fam, beams, fwdK, bwdK) in CAPTURES.items():
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr = rows[0]; col = {h: i for i, h in enumerate(hdr)}
    per = {}
    with open(os.path.join(out, f"{tag}_{name}_ncu.md"), "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none, one Newton iteration of the `{fam}` kernels, {beams} beams x {CELLS} cells (Newmark)\n\n")
        for r in rows[2:]:
            kn = r[col["Kernel Name"]].split("(")[0].split("<")[0].replace("void ", "").strip()
            f.write(f"## launch: {r[col['Kernel Name']][:90]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for w in WANT:
                if w in col:
                    f.write(f"| {w} | {r[col[w]]} | {rows[1][col[w]]} |\n")
            st = []
            for h, i in col.items():
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    try: st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                    except ValueError: pass
            f.write("\nWarp stalls per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:8]) + "\n\n")
            gb = lambda k: float(r[col[k]]) * UNIT.get(rows[1][col[k]], 1)
            per[kn] = {"dram": gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum"), "ms": float(r[col["gpu__time_duration.sum"]]) * {"us": 1e-3, "ms": 1, "ns": 1e-6}.get(rows[1][col["gpu__time_duration.sum"]], 1e-3)}
    flop = {k: fp64_flop(rep, k) for k in fwdK + bwdK}
    traffic[fam] = {"cells_per_launch": beams * CELLS, "beams": beams,
                    "forward_dram_bytes": sum(per[k]["dram"] for k in fwdK if k in per), "backward_dram_bytes": sum(per[k]["dram"] for k in bwdK if k in per),
                    "forward_fp64_flop": sum(flop[k] for k in fwdK), "backward_fp64_flop": sum(flop[k] for k in bwdK),
                    "ncu_ms": {k: v["ms"] for k, v in per.items()}}
    t = traffic[fam]
    print(f"{name}: {(t['forward_dram_bytes'] + t['backward_dram_bytes']) / t['cells_per_launch']:.0f} B/update, "
          f"{(t['forward_fp64_flop'] + t['backward_fp64_flop']) / t['cells_per_launch']:.0f} FLOP/update, ms {t['ncu_ms']}")
json.dump(traffic, open(os.path.join(out, "traffic.json"), "w"), indent=1)
