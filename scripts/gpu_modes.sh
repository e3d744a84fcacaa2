#!/bin/bash
# Runs on the GPU box: time the launch modes of the Newton iteration (pipelined / fused / split).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$MODE: value %.3f G/s kernel_ms %.3f (fwd %.3f bwd %.3f) frac %.4f'%(d['value']/1e9, r['kernel_ms'], r['fwd_ms'], r['bwd_ms'], r['frac']))"; }
MODE=fused run "$@"
MODE=pipelined BEAMGPU_PIPELINE=1 run "$@"
MODE=split BEAMGPU_SPLIT_KERNELS=1 run "$@"
