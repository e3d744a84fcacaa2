#!/bin/bash
# Runs on the GPU box: the whole GPU test suite, then the three bench configurations and the 8192-beam bundle a GPU owns at N = 8.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for a in "--config bundle65536" "--config bundle4096" "--config long" "--beams 8192"; do
  n=$(echo $a | tr -d ' -')
  timeout 300 python bench.py $a $BENCH_EXTRA > gpurun_out/bench_$n.json 2> gpurun_out/bench_$n.err; echo "bench $a rc=$?"; tail -2 gpurun_out/bench_$n.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    if d.get("impl") != "ours": continue
    r = d["roofline"]
    print("%s beams %d value %.4g e2e %.4g ms/step %.4f fwd %.4f bwd %.4f frac %.4f share %.4f traffic %s cpu %s parity %s" % (d["config"]["config"], d["config"]["beams_total"], d["value"], d["e2e"]["value"] if d["e2e"] else 0, d["ms_per_step"], r["fwd_ms"], r["bwd_ms"], r["frac"], r["kernel_share_of_step"], r["traffic"], d["cpu_baseline"]["value"] if d["cpu_baseline"] else None, d["parity"]))
PY
