#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
BEAMGPU_SPLIT_KERNELS=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err; echo rc=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_split.json'))
print(d['roofline'], d['value'])
PY
