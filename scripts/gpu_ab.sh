#!/bin/bash
# Runs on the GPU box: A/B of two prebuilt libraries (libbeamgpu_old.so, libbeamgpu_new.so), alternating, same box.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
run() { python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1: value %.3f G/s kernel_ms %.3f fwd %.3f bwd %.3f mhz %s'%(d['value']/1e9, r['kernel_ms'], r['fwd_ms'], r['bwd_ms'], d['clocks']['sm_mhz']))"; }
for rep in 1 2 3; do
  for v in old new; do cp beamfoam_b200/libbeamgpu_$v.so beamfoam_b200/libbeamgpu.so; run $v; done
done
