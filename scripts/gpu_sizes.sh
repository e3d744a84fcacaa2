#!/bin/bash
# Runs on the GPU box: split-kernel timings at several bundle sizes (sweep latency vs occupancy).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for nb in "$@"; do
  BEAMGPU_SPLIT_KERNELS=1 python bench.py --beams $nb --steps 4 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('beams $nb: fwd %.3f ms bwd %.3f ms value %.3f G/s'%(r['fwd_ms'], r['bwd_ms'], d['value']/1e9))"
done
