#!/bin/bash
# Runs on the GPU box: ncu --set full of chosen kernels of one family at one size.
# usage: gpu_prof_one.sh name mode beams kernel-regex count
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export BEAMGPU_NO_PIPELINE=1
MODES=$2 python scripts/gpu_sizes.py $3 > gpurun_out/plain_$1.log 2>&1 &&
MODES=$2 ncu --set full --clock-control none --import-source on -k "regex:$4" -s $((5 * $5)) -c $5 -f -o gpurun_out/prof_$1 python scripts/gpu_sizes.py $3 > gpurun_out/ncu_$1.log 2>&1
echo "$1 rc=$?"; cat gpurun_out/plain_$1.log
