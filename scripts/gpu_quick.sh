#!/bin/bash
# Runs on the GPU box: a fast parity subset, then fused and split timings of the bench workload.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
make -C oracle > gpurun_out/make.log 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden_and_edges.py -m gpu -q -x 2>&1 | tail -5
for mode in fused split; do
  if [ $mode = split ]; then export BEAMGPU_SPLIT_KERNELS=1; fi
  python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/quick_$mode.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$mode value %.3f G/s kernel_ms %.3f fwd %.3f bwd %.3f iters %.2f clocks %s'%(d['value']/1e9, r['kernel_ms'], r['fwd_ms'], r['bwd_ms'], d['newton_iterations_per_step'], d['clocks']))"
done
