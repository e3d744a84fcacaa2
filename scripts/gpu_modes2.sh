#!/bin/bash
# Runs on the GPU box: the launch modes side by side on the bench workload, twice.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
run() { python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1: value %.3f G/s ms/step %.3f kernel_ms %.3f fwd %.3f bwd %.3f iters %.2f mhz %s'%(d['value']/1e9, d['ms_per_step'], r['kernel_ms'], r['fwd_ms'], r['bwd_ms'], d['newton_iterations_per_step'], d['clocks']['sm_mhz']))"; }
for rep in 1 2; do
run split
BEAMGPU_FUSED_KERNEL=1 run fused
BEAMGPU_PIPELINE=1 BEAMGPU_PIPELINE_POLICY=1 run pipe_policy1
BEAMGPU_PIPELINE=1 BEAMGPU_PIPELINE_POLICY=0 run pipe_policy0
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
BEAMGPU_PIPELINE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_and_edges.py -m gpu -q -x 2>&1 | tail -2
