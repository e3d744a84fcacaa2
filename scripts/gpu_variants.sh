#!/bin/bash
# Runs on the GPU box: rebuild libbeamgpu.so with several -D configurations and time each.
# usage: gpu_variants.sh "-DBEAM_THREADS=64 -DBEAM_MIN_BLOCKS=4" "-DBEAM_PREFETCH_L1" ...
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in "$@"; do
  BEAMGPU_NVCC_EXTRA="$v" python -m beamfoam_b200.build --force > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  grep -A2 "beam_newtonILb1" beamfoam_b200/csrc/ptxas.log | tail -2 | tr '\n' ' '
  python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('| [$v] value %.3f G/s kernel_ms %.3f'%(d['value']/1e9, d['roofline']['kernel_ms']))"
done
