#!/bin/bash
# Runs on the GPU box: rebuild libbeamgpu.so with several launch configurations and time each.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for v in "$@"; do
  IFS=, read T M <<< "$v"
  BEAMGPU_NVCC_EXTRA="-DBEAM_THREADS=$T -DBEAM_MIN_BLOCKS=$M" python -m beamfoam_b200.build --force > /dev/null 2>&1
  grep -A2 "beam_newtonILb1" beamfoam_b200/csrc/ptxas.log | tail -2 | tr '\n' ' '
  python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('| T=$T M=$M value %.3f G/s kernel_ms %.3f'%(d['value']/1e9, d['roofline']['kernel_ms']))"
done
