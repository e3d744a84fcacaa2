#!/bin/bash
# Runs on the GPU box: rebuild libbeamgpu.so with several -D configurations and time each (fused and split), twice.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for rep in 1 2; do
for v in "$@"; do
  BEAMGPU_NVCC_EXTRA="$v" python -m beamfoam_b200.build --force > /dev/null 2>&1 || { echo "build failed: $v"; continue; }
  for mode in 1; do
  BEAMGPU_SPLIT_KERNELS=$mode python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('[$v] split=$mode value %.3f G/s kernel_ms %.3f fwd %.3f bwd %.3f mhz %s'%(d['value']/1e9, r['kernel_ms'], r['fwd_ms'], r['bwd_ms'], d['clocks']['sm_mhz']))"
  done
done
done
