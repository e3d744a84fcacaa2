#!/bin/bash
# Runs on the GPU box: smoke + bench (ours, then the CPU reference arm).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
nproc
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err
cat gpurun_out/bench.json
