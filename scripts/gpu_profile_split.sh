#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BEAMGPU_SPLIT_KERNELS=1
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_split.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:beam_(forward|backward)" -s 4 -c 2 -f -o gpurun_out/prof_split $CMD > gpurun_out/ncu_split.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_split.log
